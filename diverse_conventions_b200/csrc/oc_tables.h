// oc_tables.h — host-side construction of the static layout tables from ocb_config.
// Shared by the C ABI (ocb_api.cu) and the CPU emulation harness of the tests.
#pragma once
#include <stdio.h>
#include <string.h>

#include "oc_core.cuh"
#include "ocb.h"

namespace ocb {

#define OCB_TABLE_FAIL(code, ...)            \
    do {                                     \
        snprintf(err, errlen, __VA_ARGS__);  \
        return (code);                       \
    } while (0)

// tmpl receives the static part of one (W,H,C) observation plane, S*C bytes
inline int build_tables(const ocb_config* cfg, Tables* tb, uint8_t* tmpl, char* err, size_t errlen) {

    if (cfg == nullptr) OCB_TABLE_FAIL(OCB_ERR_INVALID_ARG, "config is NULL");
    if (cfg->struct_size != sizeof(ocb_config))
        OCB_TABLE_FAIL(OCB_ERR_INVALID_ARG, "ocb_config.struct_size %u != %zu (ABI mismatch)", cfg->struct_size,
                    sizeof(ocb_config));
    const int W = cfg->width, H = cfg->height, P = cfg->num_players;
    if (W < 1 || H < 1 || (long long)W * H > OCB_MAX_CELLS)
        OCB_TABLE_FAIL(OCB_ERR_BAD_LAYOUT, "grid %dx%d not in 1..%d cells", W, H, OCB_MAX_CELLS);
    if (P < 1 || P > OCB_MAX_PLAYERS) OCB_TABLE_FAIL(OCB_ERR_BAD_LAYOUT, "num_players %d not in 1..%d", P, OCB_MAX_PLAYERS);
    if (cfg->horizon < 1) OCB_TABLE_FAIL(OCB_ERR_INVALID_ARG, "horizon must be >= 1");
    const int S = W * H, C = 5 * P + 10;
    memset(tb, 0, sizeof(*tb));
    tb->W = W, tb->H = H, tb->S = S, tb->P = P, tb->C = C, tb->SC = S * C;
    tb->horizon = cfg->horizon;
    tb->rew_place = cfg->placement_in_pot_rew;
    tb->rew_dish = cfg->dish_pickup_rew;
    tb->rew_soup = cfg->soup_pickup_rew;
    for (int i = 0; i < OCB_NUM_RECIPES; ++i) {
        if (cfg->recipe_times[i] < 0 || cfg->recipe_times[i] > OCB_MAX_COOK_TIME)
            OCB_TABLE_FAIL(OCB_ERR_BAD_LAYOUT, "recipe_times[%d]=%d not in 0..%d", i, cfg->recipe_times[i], OCB_MAX_COOK_TIME);
        tb->rtime[i] = (uint8_t)cfg->recipe_times[i];
        tb->rvalue[i] = cfg->recipe_values[i];
    }
    if (W > 127) OCB_TABLE_FAIL(OCB_ERR_BAD_LAYOUT, "width %d > 127", W);
    tb->dpack = (uint32_t)(uint8_t)(int8_t)(-W) | ((uint32_t)(uint8_t)(int8_t)W << 8) | (1u << 16) | (0xFFu << 24);
    tb->uniform_time = cfg->recipe_times[0];
    for (int i = 1; i < OCB_NUM_RECIPES; ++i)
        if (cfg->recipe_times[i] != cfg->recipe_times[0]) tb->uniform_time = -1;
    int n_counters = 0;
    for (int pos = 0; pos < S; ++pos) {
        const int t = cfg->terrain[pos];
        if (t > T_TOMATO_SRC) OCB_TABLE_FAIL(OCB_ERR_BAD_LAYOUT, "terrain[%d]=%d is not a terrain code", pos, t);
        const int x = pos % W, y = pos / W;
        if ((x == 0 || y == 0 || x == W - 1 || y == H - 1) && t == T_AIR)
            OCB_TABLE_FAIL(OCB_ERR_BAD_LAYOUT, "border cell (%d,%d) is walkable", x, y);
        tb->cell_info[pos] = info_make(t, (x * H + y) * C, pos);
        if (t == T_COUNTER) tb->objcells[n_counters++] = (uint16_t)pos;
    }
    tb->n_objcells = n_counters;
    for (int pos = 0; pos < S; ++pos)
        if (cfg->terrain[pos] == T_POT) {
            if (tb->n_pots == kMaxPots) OCB_TABLE_FAIL(OCB_ERR_BAD_LAYOUT, "more than %d pots", kMaxPots);
            tb->pot_info[tb->n_pots++] = tb->cell_info[pos];
            tb->objcells[tb->n_objcells++] = (uint16_t)pos;
        }
    for (int i = 0; i < P; ++i) {
        const int x = cfg->start_player_x[i], y = cfg->start_player_y[i];
        if (x < 0 || x >= W || y < 0 || y >= H || cfg->terrain[y * W + x] != T_AIR)
            OCB_TABLE_FAIL(OCB_ERR_BAD_LAYOUT, "player %d start (%d,%d) is not a walkable cell", i, x, y);
        tb->start_pos[i] = y * W + x;
        for (int j = 0; j < i; ++j)  // the MDP never puts two players on one cell (reimplement.py:356-366); the plane update assumes it
            if (tb->start_pos[j] == tb->start_pos[i])
                OCB_TABLE_FAIL(OCB_ERR_BAD_LAYOUT, "players %d and %d start on the same cell (%d,%d)", j, i, x, y);
    }
    // static part of the encoding (setup_base_observation, reimplement.py:165-171) in (W,H,C) order
    memset(tmpl, 0, (size_t)S * C);
    for (int pos = 0; pos < S; ++pos) {
        const int t = cfg->terrain[pos];
        if (t > T_AIR) tmpl[info_slot(tb->cell_info[pos]) + t - 1 + 5 * P] = 1;
    }
    return OCB_OK;
}

#undef OCB_TABLE_FAIL
}  // namespace ocb
