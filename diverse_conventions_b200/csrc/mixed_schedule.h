// mixed_schedule.h — the pure index logic of the mixed-play collection, shared by the kernels (mixed_kernels.cu) and
// by the host-side emulation harness of the CPU tests (tests/emu/oc_emu.cpp): which worlds are forced to the main
// policy at step s, which (seat, world) a record item of step s is and which buffer slot it lands in, and the draw of
// the per-row mask.  Reference: XDPlayer.next_mp_step / MixedAgent.get_action (train/XD/xd_player.py:298-305,
// train/partner_agents.py:167-173), collect_mp_episode (xd_player.py:244-281), SharedReplayBuffer.diaginsert /
// partinsert (train/MAPPO/utils/shared_buffer.py:166,206).
#pragma once
#include <stdint.h>

#include "oc_core.cuh"

namespace ocb {

constexpr uint32_t kMixTag = 0x4D495845u;  // "MIXE": 4th Philox counter word of the mask stream

// is world j (0..L-2) of a replica forced to the main policy at step s (0..2L-1)?
OCB_HD bool mix_forced_main(int L, int s, int j) {
    const int G = L - 1;
    return s < L ? (s > 0 && j >= G - s) : (j < s - L);
}

// first recorded world of a replica and how many are recorded at step s (exactly the forced ones)
OCB_HD void mix_recorded_range(int L, int s, int* j0, int* cnt) {
    const int G = L - 1;
    if (s < L) *j0 = G - s, *cnt = s;  // s <= G
    else *j0 = 0, *cnt = s - L;
}

// buffer slot of the record of (step s, world j); only meaningful for forced worlds
OCB_HD int mix_record_slot(int L, int s, int j) { return s < L ? j - (L - 1) + s : s - L; }

// record item -> (seat, world, slot); items of a step are numbered seat-major, then replica, then world
OCB_HD bool mix_record_item(int L, int s, int N, int P, long long item, int* seat, int* world, int* slot) {
    int j0, cnt;
    mix_recorded_range(L, s, &j0, &cnt);
    const int G = L - 1, R = N / G;
    const long long per_seat = (long long)R * cnt;
    if (cnt == 0 || item >= per_seat * P) return false;
    *seat = (int)(item / per_seat);
    const long long k = item - (long long)*seat * per_seat;
    const int rep = (int)(k / cnt), j = j0 + (int)(k - (long long)rep * cnt);
    *world = rep * G + j;
    *slot = mix_record_slot(L, s, j);
    return true;
}

// the mask stream: true = the partner policy acts on agent row `row` at global env step `step` (before forcing)
OCB_HD bool mix_draw_partner(unsigned long long mix_seed, uint32_t row, unsigned long long step) {
    uint32_t r[4] = {row, (uint32_t)step, (uint32_t)(step >> 32), kMixTag};
    philox4x32_10(r, (uint32_t)mix_seed, (uint32_t)(mix_seed >> 32));
    return r[0] < 0x80000000u;
}

}  // namespace ocb
