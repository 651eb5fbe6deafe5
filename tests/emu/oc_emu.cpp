// oc_emu.cpp — CPU emulation of oc_rollout_kernel<P,G> (TEST INFRASTRUCTURE).
//
// Compiles the SAME per-lane code the sm_100a kernel runs (csrc/oc_core.cuh:
// step_world, obs_phase1/2, ActionRng, the packed object format) with g++ and
// drives it with the kernel's orchestration — a tile of WPW = 32/G worlds, G lanes
// per world executed one after another, the two plane phases separated exactly where
// the kernel has its warp barriers.  It lets the `-m "not gpu"` suite check the
// kernel logic against the oracle without a GPU.  It is not a product path and is
// not a CPU fallback: nothing in diverse_conventions_b200/ loads it.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "oc_core.cuh"
#include "oc_tables.h"

using namespace ocb;

namespace {

size_t a16(size_t x) { return (x + 15) & ~(size_t)15; }

bool g_two_halves = false;  // G code 101: one lane per world, transition as step_pre + step_post (two players)

template <int P, int G>
int rollout(const Tables& tb, const uint8_t* tmpl, int32_t* state, int N, int K, const uint8_t* actions, uint64_t seed,
            uint64_t step0, uint32_t world0, int8_t* obs, int32_t* rew, int32_t* done, uint8_t* actions_out) {
    constexpr int WPW = 32 / G;
    const int S = tb.S, SC = tb.SC, L = 1 + 6 * P + 4 * S;
    const Consts c = load_consts(tb);
    const int view_stride = (int)a16((size_t)WPW * SC);
    std::vector<uint8_t> planes((size_t)P * view_stride);
    std::vector<uint16_t> objs((size_t)S * WPW);
    for (int n0 = 0; n0 < N; n0 += WPW) {
        const int nvalid = (N - n0 < WPW) ? N - n0 : WPW;
        // every lane of the warp keeps its own copy of the world registers
        World<P> w[32];
        int cur_ret[32];
        ActionRng<P> rng[32];
        for (int lane = 0; lane < 32; ++lane) {
            const int wi = lane / G;
            const int nl = (n0 + wi < N) ? n0 + wi : N - 1;
            const int32_t* row = state + (size_t)nl * L;
            for (int i = 0; i < P; ++i) {
                const int32_t* pl = row + 1 + 6 * i;
                w[lane].pos[i] = pl[0];
                w[lane].slot[i] = info_slot(tb.cell_info[pl[0]]);
                w[lane].orient[i] = pl[1];
                w[lane].held[i] = pl[2] ? obj_make(pl[2], pl[3], pl[4], pl[5]) : 0u;
            }
            w[lane].timestep = row[0];
            cur_ret[lane] = 0;
            for (int c = 0; c < S; ++c) {
                const int32_t* oc = row + 1 + 6 * P + 4 * c;
                objs[(size_t)c * WPW + wi] = (uint16_t)(oc[0] ? obj_make(oc[0], oc[1], oc[2], oc[3]) : 0u);
            }
        }
        for (int lane = 0; lane < 32; ++lane) {
            const int wi = lane / G;
            int cd = 0, np = 0;
            for (int idx = 0; idx < tb.n_objcells; ++idx) {
                const uint32_t ci = tb.cell_info[tb.objcells[idx]];
                const uint32_t o = objs[(size_t)info_cell(ci) * WPW + wi];
                cd += (info_terrain(ci) == T_COUNTER && obj_name(o) == O_DISH);
                np += (info_terrain(ci) == T_POT) ? pot_counts(o) : 0;
            }
            w[lane].counter_dishes = cd;
            w[lane].nonempty_pots = np;
            const int nl = (n0 + wi < N) ? n0 + wi : N - 1;
            if (actions == nullptr && (step0 % ActionRng<P>::kStepsPerBlock) != 0)
                rng[lane].refill(seed, world0 + (uint32_t)nl, step0);
        }
        uint64_t t = step0;
        for (int k = 0; k < K; ++k, ++t) {
            int oldslot[32][P];
            uint32_t dirty[32][P];
            uint32_t ticked[32] = {};
            bool full[32];
            // transition: the G lanes of a world run in lockstep -> emulate by letting only the
            // first lane of each world touch the shared object array and copying its registers
            for (int wi = 0; wi < WPW; ++wi) {
                const int lane = wi * G;
                const int n = n0 + wi;
                const int nl = n < N ? n : N - 1;
                int act[P];
                if (actions == nullptr) {
                    if ((t % ActionRng<P>::kStepsPerBlock) == 0) rng[lane].refill(seed, world0 + (uint32_t)nl, t);
                    for (int i = 0; i < P; ++i) act[i] = rng[lane].action(t, i, 6);
                } else {
                    for (int i = 0; i < P; ++i) {
                        const int a = actions[((size_t)k * P + i) * N + nl];
                        act[i] = (a >= 0 && a <= 5) ? a : A_STAY;
                    }
                }
                if (actions_out && n < N)
                    for (int i = 0; i < P; ++i) actions_out[((size_t)k * P + i) * N + n] = (uint8_t)act[i];
                for (int i = 0; i < P; ++i) oldslot[lane][i] = w[lane].slot[i];
                int r;
                if constexpr (P == 2) {
                    if (g_two_halves) {  // the fused rollout's env warps: step_pre while the policy runs, step_post on the actions
                        StepPre2 pre;
                        step_pre(tb, c, w[lane], objs.data() + wi, WPW, pre);
                        r = step_post(tb, c, w[lane], objs.data() + wi, WPW, act, pre, dirty[lane], ticked[lane]);
                    } else {
                        r = step_world<P>(tb, c, w[lane], objs.data() + wi, WPW, act, dirty[lane], ticked[lane]);
                    }
                } else {
                    r = step_world<P>(tb, c, w[lane], objs.data() + wi, WPW, act, dirty[lane], ticked[lane]);
                }
                const bool d = w[lane].timestep >= c.horizon;
                cur_ret[lane] += r;
                if (d) {
                    cur_ret[lane] = 0;
                    reset_world<P>(tb, w[lane]);
                    for (int idx = 0; idx < tb.n_objcells; ++idx) objs[(size_t)tb.objcells[idx] * WPW + wi] = 0;
                }
                full[lane] = (k == 0) || d;
                for (int g = 1; g < G; ++g) {
                    w[lane + g] = w[lane];
                    rng[lane + g] = rng[lane];
                    full[lane + g] = full[lane];
                    for (int i = 0; i < P; ++i) oldslot[lane + g][i] = oldslot[lane][i], dirty[lane + g][i] = dirty[lane][i];
                    ticked[lane + g] = ticked[lane];
                }
                if (n < N) {
                    if (rew)
                        for (int i = 0; i < P; ++i) rew[((size_t)k * P + i) * N + n] = r;
                    if (done) done[(size_t)k * N + n] = d ? 1 : 0;
                }
            }
            if (obs) {
                for (int lane = 0; lane < 32; ++lane)  // phase 1, then the kernel's __syncwarp()
                    obs_phase1<P, G>(tb, planes.data() + (lane / G) * SC, view_stride, tmpl, full[lane], lane % G,
                                     oldslot[lane]);
                for (int lane = 0; lane < 32; ++lane)  // phase 2
                    obs_phase2<P, G>(tb, c, planes.data() + (lane / G) * SC, view_stride, objs.data() + lane / G, WPW,
                                     full[lane], lane % G, w[lane], dirty[lane], ticked[lane]);
                for (int v = 0; v < P; ++v)
                    memcpy(obs + (((size_t)k * P + v) * N + n0) * SC, planes.data() + (size_t)v * view_stride,
                           (size_t)nvalid * SC);
            }
        }
        // store state back
        for (int wi = 0; wi < nvalid; ++wi) {
            const int lane = wi * G;
            int32_t* row = state + (size_t)(n0 + wi) * L;
            row[0] = w[lane].timestep;
            for (int i = 0; i < P; ++i) {
                const uint32_t h = w[lane].held[i];
                int32_t* pl = row + 1 + 6 * i;
                pl[0] = w[lane].pos[i], pl[1] = w[lane].orient[i];
                pl[2] = obj_name(h), pl[3] = obj_onions(h), pl[4] = obj_tomatoes(h), pl[5] = h ? obj_tickp1(h) - 1 : 0;
            }
            for (int c = 0; c < S; ++c) {
                const uint32_t o = objs[(size_t)c * WPW + wi];
                int32_t* oc = row + 1 + 6 * P + 4 * c;
                oc[0] = obj_name(o), oc[1] = obj_onions(o), oc[2] = obj_tomatoes(o), oc[3] = o ? obj_tickp1(o) - 1 : 0;
            }
        }
    }
    return 0;
}

template <int P>
int rollout_p(int G, const Tables& tb, const uint8_t* tmpl, int32_t* state, int N, int K, const uint8_t* actions,
              uint64_t seed, uint64_t step0, uint32_t world0, int8_t* obs, int32_t* rew, int32_t* done,
              uint8_t* actions_out) {
    g_two_halves = (G == 101);
    if (G == 101) G = 1;
    switch (G) {
        case 1: return rollout<P, 1>(tb, tmpl, state, N, K, actions, seed, step0, world0, obs, rew, done, actions_out);
        case 2: return rollout<P, 2>(tb, tmpl, state, N, K, actions, seed, step0, world0, obs, rew, done, actions_out);
        case 4: return rollout<P, 4>(tb, tmpl, state, N, K, actions, seed, step0, world0, obs, rew, done, actions_out);
        case 8: return rollout<P, 8>(tb, tmpl, state, N, K, actions, seed, step0, world0, obs, rew, done, actions_out);
        default: return -1;
    }
}

}  // namespace

// state: packed int32 [N, L] (include/ocb.h), updated in place; actions uint8 [K,P,N] or
// NULL for the RNG stream; outputs as in ocb_rollout_* (any may be NULL).
extern "C" int ocemu_rollout(const ocb_config* cfg, int G, int32_t* state, int N, int K, const uint8_t* actions,
                             uint64_t seed, uint64_t step0, uint32_t world0, int8_t* obs, int32_t* rew, int32_t* done,
                             uint8_t* actions_out) {
    Tables tb;
    std::vector<uint8_t> tmpl(OCB_MAX_CELLS * (5 * OCB_MAX_PLAYERS + 10));
    char err[400];
    const int rc = build_tables(cfg, &tb, tmpl.data(), err, sizeof(err));
    if (rc != OCB_OK) return rc;
    switch (tb.P) {
        case 1: return rollout_p<1>(G, tb, tmpl.data(), state, N, K, actions, seed, step0, world0, obs, rew, done, actions_out);
        case 2: return rollout_p<2>(G, tb, tmpl.data(), state, N, K, actions, seed, step0, world0, obs, rew, done, actions_out);
        case 3: return rollout_p<3>(G, tb, tmpl.data(), state, N, K, actions, seed, step0, world0, obs, rew, done, actions_out);
        case 4: return rollout_p<4>(G, tb, tmpl.data(), state, N, K, actions, seed, step0, world0, obs, rew, done, actions_out);
    }
    return -1;
}

// ---------------------------------------------------------------------------------------------------------------
// mixed-play schedule / mask stream (csrc/mixed_schedule.h): the exact functions mix_select_kernel and
// mix_record_kernel call, exported for the CPU test against oracle/mixed_oracle.py
#include "mixed_schedule.h"

extern "C" int ocemu_mix_forced(int L, int s, int j) { return ocb::mix_forced_main(L, s, j) ? 1 : 0; }

// enumerates the record items of step s exactly as launch_mix_record / mix_record_kernel do; out = [items][3] (seat, world, slot)
extern "C" int ocemu_mix_items(int L, int s, int N, int P, int* out, int max_items) {
    const int G = L - 1, R = N / G;
    const int cnt = s < L ? s : s - L;
    const long long items = (long long)R * cnt * P;
    int n = 0;
    for (long long item = 0; item < items + 8; ++item) {  // a few past the end: the kernel's tail warps must bail out
        int seat, world, slot;
        if (!ocb::mix_record_item(L, s, N, P, item, &seat, &world, &slot)) continue;
        if (n < max_items) out[3 * n] = seat, out[3 * n + 1] = world, out[3 * n + 2] = slot;
        ++n;
    }
    return n;
}

extern "C" int ocemu_mix_draw(unsigned long long seed, unsigned int row, unsigned long long step) {
    return ocb::mix_draw_partner(seed, row, step) ? 1 : 0;
}
