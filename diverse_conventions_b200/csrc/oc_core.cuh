// oc_core.cuh — per-world Overcooked logic shared by the sm_100a kernels
// (oc_kernels.cu) and by the host-side emulation harness used in the CPU tests
// (tests/emu/oc_emu.cpp).  Everything here is a pure function of registers and of
// a few small shared-memory tables; no global memory, no synchronisation.
//
// Semantics follow the reference's Python MDP (envs/overcooked2_reimplement.py,
// "R:" below) — not its Madrona ECS systems.  The design differs from both:
//   * world state is packed: one 32-bit word per player, one 16-bit word per cell;
//   * everything the shaped reward needs (dishes lying on counters, non-empty pots) is
//     kept incrementally in registers instead of re-scanning the grid every step;
//   * one table word per cell (terrain | plane offset | cell index) serves the
//     transition and the encoder, so a step costs a handful of shared-memory loads;
//   * the observation is not re-encoded from scratch: each world keeps its
//     (W,H,C) planes resident in shared memory and only the bytes touched by the
//     transition are rewritten before the planes are streamed out.
#pragma once
#include <stdint.h>

#include "ocb.h"

#if defined(__CUDACC__)
#define OCB_HD __host__ __device__ __forceinline__
#define OCB_HDM __host__ __device__ __forceinline__
#else
#define OCB_HD static inline
#define OCB_HDM inline
#endif

namespace ocb {

constexpr int kMaxCells = OCB_MAX_CELLS;
constexpr int kMaxPlayers = OCB_MAX_PLAYERS;
constexpr int kMaxPots = 32;

// terrain (R:12-19), objects (R:5-9), actions (R:35-43)
enum : int { T_AIR = 0, T_POT = 1, T_COUNTER = 2, T_ONION_SRC = 3, T_DISH_SRC = 4, T_SERVING = 5, T_TOMATO_SRC = 6 };
enum : int { O_NONE = 0, O_TOMATO = 1, O_ONION = 2, O_DISH = 3, O_SOUP = 4 };
enum : int { A_NORTH = 0, A_SOUTH = 1, A_EAST = 2, A_WEST = 3, A_STAY = 4, A_INTERACT = 5 };

// cell info word: bits 0-2 terrain | 3-15 byte offset of the cell inside a (W,H,C)
// plane, (x*H+y)*C | 16-23 cell index (pos = y*W+x)
OCB_HD uint32_t info_make(int terrain, int slot, int cell) {
    return (uint32_t)terrain | ((uint32_t)slot << 3) | ((uint32_t)cell << 16);
}
OCB_HD int info_terrain(uint32_t ci) { return (int)(ci & 7u); }
OCB_HD int info_slot(uint32_t ci) { return (int)((ci >> 3) & 0x1FFFu); }
OCB_HD int info_cell(uint32_t ci) { return (int)(ci >> 16); }

// Static per-layout tables, built on the host (oc_tables.h: build_tables) and staged
// into shared memory by every CTA.
struct alignas(16) Tables {
    int32_t W, H, S, P, C, SC;  // SC = S*C bytes of one agent's observation
    int32_t horizon, n_pots, n_objcells;
    int32_t rew_place, rew_dish, rew_soup;
    int32_t uniform_time;  // cook time if all 16 recipes share it, else -1
    uint32_t dpack;        // int8 deltas of NORTH,SOUTH,EAST,WEST packed in one word
    int32_t start_pos[kMaxPlayers];
    int32_t rvalue[OCB_NUM_RECIPES];
    uint8_t rtime[OCB_NUM_RECIPES];
    uint32_t cell_info[kMaxCells];  // indexed by pos
    uint32_t pot_info[kMaxPots];    // info words of the POT cells
    uint16_t objcells[kMaxCells];   // cells that can hold an object: counters, then pots
};

// per-thread copies of the scalars the step needs (registers; shared-memory reads
// cannot be cached by the compiler across the warp barriers of the rollout loop)
struct Consts {
    int W, n_pots, n_objcells, horizon;
    int rew_place, rew_dish, rew_soup;
    int utime;
    uint32_t dpack, pot0, pot1;
};
OCB_HD Consts load_consts(const Tables& tb) {
    Consts c;
    c.W = tb.W, c.n_pots = tb.n_pots, c.n_objcells = tb.n_objcells, c.horizon = tb.horizon;
    c.rew_place = tb.rew_place, c.rew_dish = tb.rew_dish, c.rew_soup = tb.rew_soup;
    c.utime = tb.uniform_time, c.dpack = tb.dpack;
    c.pot0 = tb.pot_info[0], c.pot1 = tb.pot_info[1];
    return c;
}

// ---------------------------------------------------------------- packed objects
// bits 0-2 name | 3-4 tomatoes | 5-6 onions | 8-15 cooking_tick+1 ; NONE == 0.
// (ObjectState, R:46-57; recipe index 4*onions+tomatoes == bits 3..6)
OCB_HD uint32_t obj_make(int name, int onions, int tomatoes, int tick) {
    return (uint32_t)name | ((uint32_t)tomatoes << 3) | ((uint32_t)onions << 5) | ((uint32_t)(tick + 1) << 8);
}
OCB_HD int obj_name(uint32_t o) { return (int)(o & 7u); }
OCB_HD int obj_tomatoes(uint32_t o) { return (int)((o >> 3) & 3u); }
OCB_HD int obj_onions(uint32_t o) { return (int)((o >> 5) & 3u); }
OCB_HD int obj_recipe(uint32_t o) { return (int)((o >> 3) & 15u); }
OCB_HD int obj_tickp1(uint32_t o) { return (int)((o >> 8) & 0xFFu); }
OCB_HD int obj_ingredients(uint32_t o) { return obj_onions(o) + obj_tomatoes(o); }
// a pot counts as non-empty for the dish-pickup shaping (get_pot_states, R:272-281)
OCB_HD int pot_counts(uint32_t o) { return (o != 0u && (obj_tickp1(o) >= 1 || obj_ingredients(o) < 3)) ? 1 : 0; }

// player word: bits 0-11 pos | 12-13 orientation | 16-31 held object
OCB_HD uint32_t player_pack(int pos, int orient, uint32_t held) {
    return (uint32_t)pos | ((uint32_t)orient << 12) | (held << 16);
}

template <int P>
struct World {
    int pos[P];
    int slot[P];  // plane byte offset of the cell the player stands on
    int orient[P];
    uint32_t held[P];
    int timestep;
    int counter_dishes;  // DISH objects lying on COUNTER cells
    int nonempty_pots;   // pots counted by pot_counts()
};

template <int P>
OCB_HD int sel(const int (&a)[P], int idx) {
    int r = a[0];
#pragma unroll
    for (int q = 1; q < P; ++q) r = (idx == q) ? a[q] : r;
    return r;
}
template <int P>
OCB_HD uint32_t selu(const uint32_t (&a)[P], int idx) {
    uint32_t r = a[0];
#pragma unroll
    for (int q = 1; q < P; ++q) r = (idx == q) ? a[q] : r;
    return r;
}

// move_in_direction, R:22-32, branch-free: signed byte d of dpack, 0 for STAY / INTERACT
OCB_HD int dir_delta(int d, uint32_t dpack) {
    const int v = (int)(int8_t)(dpack >> ((d & 3) << 3));
    return d < A_STAY ? v : 0;
}

OCB_HD int cook_time(const Tables& tb, const Consts& c, uint32_t o) {
    return c.utime >= 0 ? c.utime : (int)tb.rtime[obj_recipe(o)];
}
// is_cooking / is_ready, R:159-163 (tick = tickp1-1)
OCB_HD bool soup_cooking(const Tables& tb, const Consts& c, uint32_t o) {
    const int tp1 = obj_tickp1(o);
    return tp1 >= 1 && tp1 <= cook_time(tb, c, o);
}
OCB_HD bool soup_ready(const Tables& tb, const Consts& c, uint32_t o) {
    const int tp1 = obj_tickp1(o);
    return tp1 >= 1 && tp1 > cook_time(tb, c, o);
}

// returns whether the pot's soup ticked (its plane bytes change)
OCB_HD bool tick_pot(const Tables& tb, const Consts& c, uint16_t* objs, int ostride, uint32_t pot_info) {
    const int cell = info_cell(pot_info);
    const uint32_t o = objs[cell * ostride];
    const bool cooking = obj_name(o) == O_SOUP && soup_cooking(tb, c, o);
    if (cooking) objs[cell * ostride] = (uint16_t)(o + 0x100u);
    return cooking;
}

// One world transition (R:381-385): resolve_interacts -> resolve_movement ->
// step_environment_effects.  `objs[cell*ostride]` is this world's object on `cell`.
// dirty[i] receives the info word of the counter / pot cell touched by player i's
// interact (or 0xFFFFFFFF).  Returns the team reward (sum over players;
// envs/overcooked2_env.py:336).
// `ticked` receives one bit per pot (index into pot0, pot1, pot_info[2..]) whose soup ticked in this step: together with
// `dirty` these are the only cells whose dynamic plane bytes changed.
template <int P>
OCB_HD int step_world(const Tables& tb, const Consts& c, World<P>& w, uint16_t* objs, int ostride, const int (&act)[P],
                      uint32_t (&dirty)[P], uint32_t& ticked) {
    int reward = 0;
    // pot snapshot "taken once before the player loop" (R:302): the running count as of step start
    const int pots_before = w.nonempty_pots;

    // resolve_interacts: players in index order on live state (R:305-353).
    // Written as straight-line selects, not as the reference's if / else tree: a warp holds 32 different worlds, so a
    // branchy version executes every arm of the tree one after the other anyway (~1,200 cycles of a 2,200-cycle
    // transition), while here the arms are independent dataflow the scheduler overlaps, and the dependent chain is the
    // handful of operations from a player's target object to its new value.  Only the load / store of the target cell
    // are predicated.  Arm by arm the values are those of the tree (each predicate below names its reference lines).
    uint32_t fci[P];  // the cell each player faces, from the pose at step start (R:309-310: interacts never move anyone)
#pragma unroll
    for (int i = 0; i < P; ++i) fci[i] = tb.cell_info[w.pos[i] + dir_delta(w.orient[i], c.dpack)];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const bool inter = act[i] == A_INTERACT;
        const uint32_t ci = fci[i];
        const int t = info_terrain(ci);
        const int tgt = info_cell(ci);
        const uint32_t h = w.held[i];
        const int hn = obj_name(h);
        const bool is_counter = inter && t == T_COUNTER;        // R:313-319
        const bool is_pot = inter && t == T_POT && h != 0u;     // R:331-349 (empty hands do nothing at a pot)
        uint32_t o = 0u;
        if (is_counter || is_pot) o = objs[tgt * ostride];
        const int on = obj_name(o);
        const bool c_place = is_counter && h != 0u && o == 0u;
        const bool c_pick = is_counter && h == 0u && o != 0u;
        const bool p_take = is_pot && hn == O_DISH && o != 0u && soup_ready(tb, c, o);        // R:332-336
        const bool p_ingr = is_pot && (hn == O_ONION || hn == O_TOMATO);                      // R:337-349
        uint32_t o2 = (o == 0u) ? obj_make(O_SOUP, 0, 0, -1) : o;
        const bool p_add = p_ingr && !(obj_tickp1(o2) >= 1 || obj_ingredients(o2) == 3);
        o2 += p_add ? ((hn == O_ONION) ? (1u << 5) : (1u << 3)) : 0u;
        // soup_to_be_cooked_at_location (R:287-296) and full -> auto start
        o2 |= (obj_name(o2) == O_SOUP && obj_tickp1(o2) == 0 && obj_ingredients(o2) == 3) ? (1u << 8) : 0u;
        const bool free_hands = inter && h == 0u;  // the three dispensers only serve empty hands (R:320-327)
        const bool d_onion = free_hands && t == T_ONION_SRC, d_tomato = free_hands && t == T_TOMATO_SRC;
        const bool d_dish = free_hands && t == T_DISH_SRC;
        const bool serve = inter && t == T_SERVING && hn == O_SOUP;  // R:350-353, deliver_soup R:283-285

        uint32_t newo = c_place ? h : o;
        newo = (c_pick || p_take) ? 0u : newo;
        newo = p_ingr ? o2 : newo;
        if (c_place || c_pick || is_pot) objs[tgt * ostride] = (uint16_t)newo;

        reward += p_take ? c.rew_soup : 0;
        reward += p_add ? c.rew_place : 0;
        if (P == 2) {  // is_dish_pickup_useful R:261-270 (held objects as they are now: earlier players already acted)
            int held_dishes = 0;
#pragma unroll
            for (int j = 0; j < P; ++j) held_dishes += (obj_name(w.held[j]) == O_DISH);
            reward += (d_dish && w.counter_dishes == 0 && held_dishes < pots_before) ? c.rew_dish : 0;
        }
        const int served = tb.rvalue[obj_recipe(h)];  // (index 0..15 whatever h is)
        reward += serve ? served : 0;

        uint32_t nh = (c_place || p_add || serve) ? 0u : h;
        nh = (c_pick || p_take) ? o : nh;
        nh = d_onion ? obj_make(O_ONION, 0, 0, -1) : nh;
        nh = d_tomato ? obj_make(O_TOMATO, 0, 0, -1) : nh;
        nh = d_dish ? obj_make(O_DISH, 0, 0, -1) : nh;
        w.held[i] = nh;
        w.counter_dishes += (c_place && hn == O_DISH) ? 1 : 0;
        w.counter_dishes -= (c_pick && on == O_DISH) ? 1 : 0;
        w.nonempty_pots += is_pot ? pot_counts(newo) - pot_counts(o) : 0;
        dirty[i] = (is_counter || is_pot) ? ci : 0xFFFFFFFFu;
    }

    // resolve_movement R:368-371, _move_if_direction R:393-399
    int np[P], ns[P], no[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int a = act[i];
        const uint32_t ci = tb.cell_info[w.pos[i] + dir_delta(a, c.dpack)];  // INTERACT/STAY -> own cell
        const bool walk = (a < A_STAY) && (info_terrain(ci) == T_AIR);
        np[i] = walk ? info_cell(ci) : w.pos[i];
        ns[i] = walk ? info_slot(ci) : w.slot[i];
        no[i] = (a < A_STAY) ? a : w.orient[i];
    }
    // _handle_collisions R:356-366: any colliding pair freezes every position
    bool blocked = false;
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int j = i + 1; j < P; ++j)
            blocked |= (np[i] == np[j]) || (np[i] == w.pos[j] && w.pos[i] == np[j]);
#pragma unroll
    for (int i = 0; i < P; ++i) {
        w.pos[i] = blocked ? w.pos[i] : np[i];
        w.slot[i] = blocked ? w.slot[i] : ns[i];
        w.orient[i] = no[i];
    }

    // step_environment_effects R:373-379 (cooking soups only ever sit in pots)
    w.timestep += 1;
    uint32_t tk = 0;
    if (c.n_pots > 0) tk |= tick_pot(tb, c, objs, ostride, c.pot0) ? 1u : 0u;
    if (c.n_pots > 1) tk |= tick_pot(tb, c, objs, ostride, c.pot1) ? 2u : 0u;
    for (int q = 2; q < c.n_pots; ++q) tk |= tick_pot(tb, c, objs, ostride, tb.pot_info[q]) ? (1u << (q & 31)) : 0u;
    ticked = c.n_pots > 32 ? 0xFFFFFFFFu : tk;  // (more pots than bits: re-encode all of them)
    return reward;
}

// ---------------------------------------------------------------- the transition in two halves (two players)
// In the fused rollout the env warps wait most of a step for the sampled actions, and what follows the hand-off is on the
// step's critical path.  Nearly all of step_world depends on the STATE only: which cell each player faces and what an
// INTERACT would do there, where each of the four moves would lead, what the pots hold.  `step_pre` computes those
// outcomes while the policy forward runs; `step_post` picks among them once the joint action is known, resolves the
// collisions, applies the writes and ticks the pots — the same values as step_world (the emulated-kernel test and the
// fused-vs-per-step bit-identity tests run both).
struct InteractEval {  // what a player's INTERACT does on a given state
    uint32_t held, newo, dirty;  // held object / faced cell's object afterwards, info word of the touched counter / pot cell
    int reward, d_cd, d_np;      // reward, change of counter_dishes / nonempty_pots
    bool wr;                     // the faced cell is written
};
// `h` = the player's held object, `o_cell` = the object on the faced cell `ci` (anything if ci is neither counter nor pot),
// `held_dishes` / `counter_dishes` / `pots_before` as step_world sees them at this player's turn.  Same arms and
// reference lines as the loop body of step_world.
template <int P>
OCB_HD InteractEval interact_eval(const Tables& tb, const Consts& c, uint32_t ci, uint32_t h, uint32_t o_cell, int held_dishes,
                                  int counter_dishes, int pots_before) {
    const int t = info_terrain(ci);
    const int hn = obj_name(h);
    const bool is_counter = t == T_COUNTER;
    const bool is_pot = t == T_POT && h != 0u;
    const uint32_t o = (is_counter || is_pot) ? o_cell : 0u;
    const int on = obj_name(o);
    const bool c_place = is_counter && h != 0u && o == 0u;
    const bool c_pick = is_counter && h == 0u && o != 0u;
    const bool p_take = is_pot && hn == O_DISH && o != 0u && soup_ready(tb, c, o);
    const bool p_ingr = is_pot && (hn == O_ONION || hn == O_TOMATO);
    uint32_t o2 = (o == 0u) ? obj_make(O_SOUP, 0, 0, -1) : o;
    const bool p_add = p_ingr && !(obj_tickp1(o2) >= 1 || obj_ingredients(o2) == 3);
    o2 += p_add ? ((hn == O_ONION) ? (1u << 5) : (1u << 3)) : 0u;
    o2 |= (obj_name(o2) == O_SOUP && obj_tickp1(o2) == 0 && obj_ingredients(o2) == 3) ? (1u << 8) : 0u;
    const bool free_hands = h == 0u;
    const bool d_onion = free_hands && t == T_ONION_SRC, d_tomato = free_hands && t == T_TOMATO_SRC;
    const bool d_dish = free_hands && t == T_DISH_SRC;
    const bool serve = t == T_SERVING && hn == O_SOUP;
    InteractEval e;
    uint32_t newo = c_place ? h : o;
    newo = (c_pick || p_take) ? 0u : newo;
    newo = p_ingr ? o2 : newo;
    e.newo = newo;
    e.wr = c_place || c_pick || is_pot;
    int reward = (p_take ? c.rew_soup : 0) + (p_add ? c.rew_place : 0);
    if (P == 2) reward += (d_dish && counter_dishes == 0 && held_dishes < pots_before) ? c.rew_dish : 0;
    const int served = tb.rvalue[obj_recipe(h)];
    reward += serve ? served : 0;
    e.reward = reward;
    uint32_t nh = (c_place || p_add || serve) ? 0u : h;
    nh = (c_pick || p_take) ? o : nh;
    nh = d_onion ? obj_make(O_ONION, 0, 0, -1) : nh;
    nh = d_tomato ? obj_make(O_TOMATO, 0, 0, -1) : nh;
    nh = d_dish ? obj_make(O_DISH, 0, 0, -1) : nh;
    e.held = nh;
    e.d_cd = ((c_place && hn == O_DISH) ? 1 : 0) - ((c_pick && on == O_DISH) ? 1 : 0);
    e.d_np = is_pot ? pot_counts(newo) - pot_counts(o) : 0;
    e.dirty = (is_counter || is_pot) ? ci : 0xFFFFFFFFu;
    return e;
}

struct StepPre2 {
    InteractEval e0, e1a, e1b;  // player 0; player 1 if player 0 does not / does interact
    int tgt0, tgt1;             // faced cells
    uint32_t np[2];             // destination cell of moves NORTH..WEST, 8 bits each (own cell where the move is blocked)
    uint32_t ns_lo[2], ns_hi[2];  // plane offsets of those cells, 16 bits each
    uint32_t pot_o[2];          // objects in pot0 / pot1 at the start of the step
};

OCB_HD void step_pre(const Tables& tb, const Consts& c, const World<2>& w, const uint16_t* objs, int ostride, StepPre2& p) {
    const uint32_t f0 = tb.cell_info[w.pos[0] + dir_delta(w.orient[0], c.dpack)];
    const uint32_t f1 = tb.cell_info[w.pos[1] + dir_delta(w.orient[1], c.dpack)];
    p.tgt0 = info_cell(f0), p.tgt1 = info_cell(f1);
    const uint32_t o0 = objs[p.tgt0 * ostride], o1 = objs[p.tgt1 * ostride];
    const int d0 = obj_name(w.held[0]) == O_DISH, d1 = obj_name(w.held[1]) == O_DISH;
    p.e0 = interact_eval<2>(tb, c, f0, w.held[0], o0, d0 + d1, w.counter_dishes, w.nonempty_pots);
    p.e1a = interact_eval<2>(tb, c, f1, w.held[1], o1, d0 + d1, w.counter_dishes, w.nonempty_pots);
    const uint32_t o1b = (p.tgt1 == p.tgt0 && p.e0.wr) ? p.e0.newo : o1;
    p.e1b = interact_eval<2>(tb, c, f1, w.held[1], o1b, (obj_name(p.e0.held) == O_DISH) + d1, w.counter_dishes + p.e0.d_cd,
                             w.nonempty_pots);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        uint32_t cells = 0u, lo = 0u, hi = 0u;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const uint32_t ci = tb.cell_info[w.pos[i] + dir_delta(a, c.dpack)];
            const bool walk = info_terrain(ci) == T_AIR;
            const uint32_t cell = walk ? (uint32_t)info_cell(ci) : (uint32_t)w.pos[i];
            const uint32_t slot = walk ? (uint32_t)info_slot(ci) : (uint32_t)w.slot[i];
            cells |= cell << (8 * a);
            if (a < 2) lo |= slot << (16 * a); else hi |= slot << (16 * (a - 2));
        }
        p.np[i] = cells, p.ns_lo[i] = lo, p.ns_hi[i] = hi;
    }
    p.pot_o[0] = c.n_pots > 0 ? (uint32_t)objs[info_cell(c.pot0) * ostride] : 0u;
    p.pot_o[1] = c.n_pots > 1 ? (uint32_t)objs[info_cell(c.pot1) * ostride] : 0u;
}

// the second half: same results as step_world<2>(tb, c, w, objs, ostride, act, dirty, ticked) on the state `p` was computed from
OCB_HD int step_post(const Tables& tb, const Consts& c, World<2>& w, uint16_t* objs, int ostride, const int (&act)[2],
                     const StepPre2& p, uint32_t (&dirty)[2], uint32_t& ticked) {
    const bool i0 = act[0] == A_INTERACT, i1 = act[1] == A_INTERACT;
    InteractEval e1;  // (field by field: a reference picked at run time would put the outcomes in local memory)
    e1.held = i0 ? p.e1b.held : p.e1a.held, e1.newo = i0 ? p.e1b.newo : p.e1a.newo, e1.dirty = i0 ? p.e1b.dirty : p.e1a.dirty;
    e1.reward = i0 ? p.e1b.reward : p.e1a.reward, e1.d_cd = i0 ? p.e1b.d_cd : p.e1a.d_cd, e1.d_np = i0 ? p.e1b.d_np : p.e1a.d_np;
    e1.wr = i0 ? p.e1b.wr : p.e1a.wr;
    const bool w0 = i0 && p.e0.wr, w1 = i1 && e1.wr;
    const uint32_t n0 = p.e0.newo, n1 = e1.newo;
    if (w0) objs[p.tgt0 * ostride] = (uint16_t)n0;
    if (w1) objs[p.tgt1 * ostride] = (uint16_t)n1;
    const int reward = (i0 ? p.e0.reward : 0) + (i1 ? e1.reward : 0);
    w.held[0] = i0 ? p.e0.held : w.held[0];
    w.held[1] = i1 ? e1.held : w.held[1];
    w.counter_dishes += (i0 ? p.e0.d_cd : 0) + (i1 ? e1.d_cd : 0);
    w.nonempty_pots += (i0 ? p.e0.d_np : 0) + (i1 ? e1.d_np : 0);
    dirty[0] = i0 ? p.e0.dirty : 0xFFFFFFFFu;
    dirty[1] = i1 ? e1.dirty : 0xFFFFFFFFu;

    // resolve_movement / _handle_collisions (R:356-371, 393-399) from the precomputed destinations
    int np[2], ns[2], no[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int a = act[i];
        const bool mv = a < A_STAY;
        const uint32_t slots = (a & 2) ? p.ns_hi[i] : p.ns_lo[i];
        np[i] = mv ? (int)((p.np[i] >> (8 * (a & 3))) & 0xFFu) : w.pos[i];
        ns[i] = mv ? (int)((slots >> (16 * (a & 1))) & 0xFFFFu) : w.slot[i];
        no[i] = mv ? a : w.orient[i];
    }
    const bool blocked = (np[0] == np[1]) || (np[0] == w.pos[1] && w.pos[0] == np[1]);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        w.pos[i] = blocked ? w.pos[i] : np[i];
        w.slot[i] = blocked ? w.slot[i] : ns[i];
        w.orient[i] = no[i];
    }

    // step_environment_effects (R:373-379): the first two pots from the values read in step_pre, patched by this step's writes
    w.timestep += 1;
    uint32_t tk = 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (q >= c.n_pots) break;
        const int cell = info_cell(q == 0 ? c.pot0 : c.pot1);
        uint32_t o = p.pot_o[q];
        o = (w0 && p.tgt0 == cell) ? n0 : o;
        o = (w1 && p.tgt1 == cell) ? n1 : o;
        const bool cooking = obj_name(o) == O_SOUP && soup_cooking(tb, c, o);
        if (cooking) objs[cell * ostride] = (uint16_t)(o + 0x100u);
        tk |= cooking ? (1u << q) : 0u;
    }
    for (int q = 2; q < c.n_pots; ++q) tk |= tick_pot(tb, c, objs, ostride, tb.pot_info[q]) ? (1u << (q & 31)) : 0u;
    ticked = c.n_pots > 32 ? 0xFFFFFFFFu : tk;
    return reward;
}

template <int P>
OCB_HD void reset_world(const Tables& tb, World<P>& w) {  // R:387-391
#pragma unroll
    for (int i = 0; i < P; ++i) {
        w.pos[i] = tb.start_pos[i];
        w.slot[i] = info_slot(tb.cell_info[tb.start_pos[i]]);
        w.orient[i] = 0;
        w.held[i] = 0u;
    }
    w.timestep = 0;
    w.counter_dishes = 0;
    w.nonempty_pots = 0;
}

// ---------------------------------------------------------------- observation bytes
// lossless_state_encoding, R:173-259; plane layout (W,H,C) as delivered by the env
// adapter (envs/overcooked2_env.py:322-325).  shift = 5P.

// the five dynamic channels shift+5..shift+9 of a counter / pot cell (R:177-211)
template <int P>
OCB_HD void encode_cell(uint8_t* plane, uint32_t ci, uint32_t o) {
    uint8_t* px = plane + info_slot(ci) + 5 * P + 5;
    const int name = obj_name(o);
    const bool soup_in_pot = (name == O_SOUP) && (info_terrain(ci) == T_POT);
    const int tp1 = obj_tickp1(o);
    px[0] = soup_in_pot ? (uint8_t)obj_onions(o) : (uint8_t)0;
    px[1] = (soup_in_pot && tp1 >= 1) ? (uint8_t)(tp1 - 1) : (uint8_t)0;
    px[2] = (name == O_SOUP && !soup_in_pot) ? 1 : 0;
    px[3] = (name == O_DISH) ? 1 : 0;
    px[4] = (name == O_ONION) ? 1 : 0;
}

// The same five bytes as two values: byte 0 (channel shift+5) and bytes 1..4 as one little-endian word.  They do not
// depend on the viewer, so a caller that updates all views of a cell computes them once (`store_cell`).
OCB_HD void cell_bytes(uint32_t ci, uint32_t o, uint32_t& b0, uint32_t& w14) {
    const int name = obj_name(o);
    const bool soup_in_pot = (name == O_SOUP) && (info_terrain(ci) == T_POT);
    const int tp1 = obj_tickp1(o);
    b0 = soup_in_pot ? (uint32_t)obj_onions(o) : 0u;
    w14 = ((soup_in_pot && tp1 >= 1) ? (uint32_t)(tp1 - 1) : 0u) | ((name == O_SOUP && !soup_in_pot) ? (1u << 8) : 0u) |
          ((name == O_DISH) ? (1u << 16) : 0u) | ((name == O_ONION) ? (1u << 24) : 0u);
}
// P == 2: the five channels are bytes 15..19 of the cell's 20 — the last byte of one aligned word and the whole next one
template <int P>
OCB_HD void store_cell(uint8_t* plane, uint32_t ci, uint32_t b0, uint32_t w14) {
    uint8_t* px = plane + info_slot(ci) + 5 * P + 5;
    px[0] = (uint8_t)b0;
    if (P == 2) {
        *reinterpret_cast<uint32_t*>(px + 1) = w14;
    } else {
        px[1] = (uint8_t)w14, px[2] = (uint8_t)(w14 >> 8), px[3] = (uint8_t)(w14 >> 16), px[4] = (uint8_t)(w14 >> 24);
    }
}

// player i as seen by `viewer` (R:221-257): position one-hot, orientation one-hot of
// the relative player index, held object drawn on the holder's cell
template <int P>
OCB_HD void poke_player(uint8_t* plane, int viewer, int i, int slot, int orient, uint32_t held) {
    uint8_t* px = plane + slot;
    const int rel = (i == viewer) ? 0 : (i < viewer ? i + 1 : i);
    px[rel] = 1;
    px[P + 4 * rel + orient] = 1;
    const int name = obj_name(held);
    if (name >= O_ONION) px[5 * P + 11 - name] = 1;  // SOUP -> shift+7, DISH -> shift+8, ONION -> shift+9
}

// a cell a player has left: players only stand on AIR, whose static bytes are all 0
template <int P>
OCB_HD void clear_cell(uint8_t* plane, int slot) {
    constexpr int C = 5 * P + 10;
    uint8_t* px = plane + slot;
    if (C % 4 == 0) {
        uint32_t* p4 = reinterpret_cast<uint32_t*>(px);
#pragma unroll
        for (int q = 0; q < C / 4; ++q) p4[q] = 0u;
    } else {
#pragma unroll
        for (int q = 0; q < C; ++q) px[q] = 0;
    }
}

// Role-split plane maintenance.  A world is served by G lanes (g = 0..G-1); the
// caller separates phase 1 and phase 2 with a warp barrier.
//   full == true : rebuild from the static template (launch start, episode reset)
//   full == false: rewrite only what the transition touched.
// planes: this world's plane of view v is at planes + v*view_stride.
template <int P, int G>
OCB_HD void obs_phase1(const Tables& tb, uint8_t* planes, int view_stride, const uint8_t* tmpl, bool full, int g,
                       const int (&oldslot)[P]) {
    if (full) {
        if (tb.SC % 4 == 0) {
            const uint32_t* t4 = reinterpret_cast<const uint32_t*>(tmpl);
            const int n4 = tb.SC >> 2;
            for (int v = 0; v < P; ++v) {
                uint32_t* d4 = reinterpret_cast<uint32_t*>(planes + v * view_stride);
                for (int j = g; j < n4; j += G) d4[j] = t4[j];
            }
        } else {
            for (int v = 0; v < P; ++v)
                for (int j = g; j < tb.SC; j += G) planes[v * view_stride + j] = tmpl[j];
        }
    } else {
#pragma unroll
        for (int j0 = 0; j0 < P * P; j0 += G) {
            const int j = j0 + g;
            if (j < P * P) clear_cell<P>(planes + (j / P) * view_stride, sel<P>(oldslot, j % P));
        }
    }
}

template <int P, int G>
OCB_HD void obs_phase2(const Tables& tb, const Consts& c, uint8_t* planes, int view_stride, const uint16_t* objs,
                       int ostride, bool full, int g, const World<P>& w, const uint32_t (&dirty)[P], uint32_t ticked) {
    if (full) {
        for (int idx = g; idx < c.n_objcells; idx += G) {
            const uint32_t ci = tb.cell_info[tb.objcells[idx]];
            const uint32_t o = objs[info_cell(ci) * ostride];
            if (o != 0u)
                for (int v = 0; v < P; ++v) encode_cell<P>(planes + v * view_stride, ci, o);
        }
    } else if (G == 1) {
        // one lane does all views of its world: the bytes of a touched cell are computed once and stored per view
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const uint32_t ci = dirty[i];
            if (ci != 0xFFFFFFFFu) {
                uint32_t b0, w14;
                cell_bytes(ci, objs[info_cell(ci) * ostride], b0, w14);
#pragma unroll
                for (int v = 0; v < P; ++v) store_cell<P>(planes + v * view_stride, ci, b0, w14);
            }
        }
        for (int q = 0; q < c.n_pots; ++q) {
            if (!((ticked >> (q & 31)) & 1u)) continue;
            const uint32_t ci = q == 0 ? c.pot0 : q == 1 ? c.pot1 : tb.pot_info[q];
            uint32_t b0, w14;
            cell_bytes(ci, objs[info_cell(ci) * ostride], b0, w14);
#pragma unroll
            for (int v = 0; v < P; ++v) store_cell<P>(planes + v * view_stride, ci, b0, w14);
        }
    } else {
        // (interact targets + pots) x views
#pragma unroll
        for (int j0 = 0; j0 < P * P; j0 += G) {
            const int j = j0 + g;
            if (j < P * P) {
                const uint32_t ci = selu<P>(dirty, j / P);
                if (ci != 0xFFFFFFFFu) encode_cell<P>(planes + (j % P) * view_stride, ci, objs[info_cell(ci) * ostride]);
            }
        }
        // pots whose soup ticked (a pot that was interacted with is among the dirty cells above; an idle, a full-but-unstarted
        // or a finished pot keeps its bytes): with random actions pots rarely cook, and re-encoding every pot in every view
        // every step was a tenth of the per-step dependent chain
        for (int j = g; j < c.n_pots * P; j += G) {
            const int q = j / P;
            if (!((ticked >> (q & 31)) & 1u)) continue;
            const uint32_t ci = q == 0 ? c.pot0 : q == 1 ? c.pot1 : tb.pot_info[q];
            encode_cell<P>(planes + (j % P) * view_stride, ci, objs[info_cell(ci) * ostride]);
        }
    }
#pragma unroll
    for (int j0 = 0; j0 < P * P; j0 += G) {
        const int j = j0 + g;
        if (j < P * P) {
            const int v = j / P, i = j % P;
            poke_player<P>(planes + v * view_stride, v, i, sel<P>(w.slot, i), sel<P>(w.orient, i), selu<P>(w.held, i));
        }
    }
}

// ---------------------------------------------------------------- action RNG
// Philox4x32-10 (Salmon et al., SC'11), counter = (world, block_lo, block_hi, 0),
// key = (seed_lo, seed_hi).  One block serves 8/P_pad consecutive steps.
OCB_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
OCB_HD void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0, c[1] = lo1, c[2] = n2, c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
template <int P>
struct ActionRng {
    static constexpr int kPad = P <= 2 ? 2 : 4;
    static constexpr int kStepsPerBlock = 8 / kPad;
    uint32_t r[4];
    OCB_HDM void refill(uint64_t seed, uint32_t world, uint64_t step) {
        const uint64_t block = step / kStepsPerBlock;
        r[0] = world, r[1] = (uint32_t)block, r[2] = (uint32_t)(block >> 32), r[3] = 0u;
        philox4x32_10(r, (uint32_t)seed, (uint32_t)(seed >> 32));
    }
    // 16-bit slice number sub*kPad + player, mapped to 0..num_actions-1
    OCB_HDM int action(uint64_t step, int player, int num_actions) const {
        const int h = (int)(step % kStepsPerBlock) * kPad + player;
        const int wsel = h >> 1;  // two-level select on its bits (a compare chain compiles to branches)
        const uint32_t lo = (wsel & 1) ? r[1] : r[0], hi = (wsel & 1) ? r[3] : r[2];
        const uint32_t word = (wsel & 2) ? hi : lo;
        const uint32_t half = (h & 1) ? (word >> 16) : (word & 0xFFFFu);
        return (int)((half * (uint32_t)num_actions) >> 16);
    }
};

}  // namespace ocb
