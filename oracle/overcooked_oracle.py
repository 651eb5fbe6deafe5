"""CPU restatement (Python) of the reference's simplified-Overcooked MDP.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg use it, and only as the checker / the timed CPU baseline.

Parity status: PINNED.  The reference ships no golden vectors for this path
(SURVEY.md section 8c), so the oracle is pinned against outputs of the reference's own
Python env run in the build container: ``tests/golden/make_golden.py`` imports
``/root/reference`` (envs/overcooked2_env.py, envs/overcooked2_reimplement.py) unmodified
and records trajectories committed under ``tests/golden/``; ``tests/test_oracle_golden.py``
replays them through this file and through ``oracle/ocb_oracle.c``.

The restatement works on the packed per-world state of the C ABI (include/ocb.h):
    row[0] timestep
    row[1+6i ...]   player i: pos, orientation, held name/onions/tomatoes/tick
    row[1+6P+4c..]  cell c:   object name/onions/tomatoes/tick      (name 0 == NONE)
Each function cites the reference lines it follows.
"""
from __future__ import annotations

import numpy as np

# reimplement.py:4-19
NONE, TOMATO, ONION, DISH, SOUP = 0, 1, 2, 3, 4
AIR, POT, COUNTER, ONION_SOURCE, DISH_SOURCE, SERVING, TOMATO_SOURCE = 0, 1, 2, 3, 4, 5, 6
# reimplement.py:35-43
NORTH, SOUTH, EAST, WEST, STAY, INTERACT = 0, 1, 2, 3, 4, 5
MAX_INGREDIENTS = 3


class OvercookedOracle:
    """N independent worlds stepped one after another on the CPU."""

    def __init__(self, params, num_worlds: int):
        p = params.as_dict() if hasattr(params, "as_dict") else dict(params)
        self.W, self.H = int(p["width"]), int(p["height"])
        self.S = self.W * self.H
        self.P = int(p["num_players"])
        self.C = 5 * self.P + 10
        self.terrain = [int(t) for t in p["terrain"]]
        self.start = [y * self.W + x for x, y in zip(p["start_player_x"], p["start_player_y"])]
        self.horizon = int(p["horizon"])
        self.r_place = int(p["placement_in_pot_rew"])
        self.r_dish = int(p["dish_pickup_rew"])
        self.r_soup = int(p["soup_pickup_rew"])
        self.values = [int(v) for v in p["recipe_values"]]
        self.times = [int(v) for v in p["recipe_times"]]
        self.N = int(num_worlds)
        self.L = 1 + 6 * self.P + 4 * self.S
        self.cell0 = 1 + 6 * self.P
        self.state = np.zeros((self.N, self.L), dtype=np.int32)
        # static part of the encoding, reimplement.py:165-171
        self.base = np.zeros((self.S, self.C), dtype=np.int8)
        for pos, t in enumerate(self.terrain):
            if t > AIR:
                self.base[pos, t - 1 + 5 * self.P] = 1
        self.reset()

    # ------------------------------------------------------------------ state
    def start_row(self):
        """reimplement.py:387-391: start cells, facing NORTH, empty hands, no objects."""
        row = [0] * self.L
        for i in range(self.P):
            row[1 + 6 * i] = self.start[i]
        return row

    def reset(self, worlds=None):
        row = np.asarray(self.start_row(), dtype=np.int32)
        if worlds is None:
            self.state[:] = row
        else:
            self.state[worlds] = row

    def get_state(self):
        return self.state.copy()

    def set_state(self, st):
        st = np.asarray(st, dtype=np.int32).reshape(self.N, self.L)
        self.state[:] = st

    # ------------------------------------------------------------------ helpers
    def _delta(self, d):
        # reimplement.py:22-32
        return (-self.W, self.W, 1, -1, 0)[d]

    def _recipe(self, on, tom):
        # reimplement.py:56-57
        return (MAX_INGREDIENTS + 1) * on + tom

    def _cooking(self, on, tom, tick):
        # reimplement.py:159-160
        return 0 <= tick < self.times[self._recipe(on, tom)]

    def _ready(self, on, tom, tick):
        # reimplement.py:162-163
        return tick >= 0 and tick >= self.times[self._recipe(on, tom)]

    # ------------------------------------------------------------------ transition
    def step_row(self, row, acts):
        """One world transition in place; returns the per-player reward list.

        reimplement.py:381-385: resolve_interacts -> resolve_movement ->
        step_environment_effects.
        """
        P, S, T, c0 = self.P, self.S, self.terrain, self.cell0
        rew = [0] * P

        # --- resolve_interacts, reimplement.py:301-354
        # pot snapshot taken once before the player loop (:302, get_pot_states :272-281)
        non_empty_pots = 0
        for pos in range(S):
            if T[pos] == POT and row[c0 + 4 * pos] != NONE:
                o = c0 + 4 * pos
                if row[o + 3] >= 0 or row[o + 1] + row[o + 2] < MAX_INGREDIENTS:
                    non_empty_pots += 1

        for i in range(P):
            if acts[i] != INTERACT:
                continue
            pl = 1 + 6 * i
            tgt = row[pl] + self._delta(row[pl + 1])  # pre-move pose (:309-310)
            t = T[tgt]
            o = c0 + 4 * tgt
            holding = row[pl + 2] != NONE
            if t == COUNTER:  # :313-319
                if holding and row[o] == NONE:
                    row[o:o + 4] = row[pl + 2:pl + 6]
                    row[pl + 2:pl + 6] = [0, 0, 0, 0]
                elif not holding and row[o] != NONE:
                    row[pl + 2:pl + 6] = row[o:o + 4]
                    row[o:o + 4] = [0, 0, 0, 0]
            elif t == ONION_SOURCE and not holding:  # :320-321
                row[pl + 2:pl + 6] = [ONION, 0, 0, -1]
            elif t == TOMATO_SOURCE and not holding:  # :322-323
                row[pl + 2:pl + 6] = [TOMATO, 0, 0, -1]
            elif t == DISH_SOURCE and not holding:  # :324-327, is_dish_pickup_useful :261-270
                useful = False
                if P == 2:
                    held_dishes = sum(1 for j in range(P) if row[1 + 6 * j + 2] == DISH)
                    dish_on_counter = any(T[q] == COUNTER and row[c0 + 4 * q] == DISH for q in range(S))
                    useful = (not dish_on_counter) and held_dishes < non_empty_pots
                if useful:
                    rew[i] += self.r_dish
                row[pl + 2:pl + 6] = [DISH, 0, 0, -1]
            elif t == POT and holding:  # :331-349
                if row[pl + 2] == DISH and row[o] != NONE and self._ready(row[o + 1], row[o + 2], row[o + 3]):
                    row[pl + 2:pl + 6] = row[o:o + 4]  # :333-336
                    row[o:o + 4] = [0, 0, 0, 0]
                    rew[i] += self.r_soup
                elif row[pl + 2] in (ONION, TOMATO):  # :337-349
                    if row[o] == NONE:
                        row[o:o + 4] = [SOUP, 0, 0, -1]
                    if not (row[o + 3] >= 0 or row[o + 1] + row[o + 2] == MAX_INGREDIENTS):
                        if row[pl + 2] == ONION:
                            row[o + 1] += 1
                        else:
                            row[o + 2] += 1
                        row[pl + 2:pl + 6] = [0, 0, 0, 0]
                        rew[i] += self.r_place
                    n_ing = row[o + 1] + row[o + 2]
                    idle = (row[o] == SOUP and not self._cooking(row[o + 1], row[o + 2], row[o + 3])
                            and not self._ready(row[o + 1], row[o + 2], row[o + 3]) and n_ing > 0)
                    if idle and n_ing == MAX_INGREDIENTS:  # :348-349 (auto start)
                        row[o + 3] = 0
            elif t == SERVING and holding:  # :350-353
                if row[pl + 2] == SOUP:
                    rew[i] += self.values[self._recipe(row[pl + 3], row[pl + 4])]
                    row[pl + 2:pl + 6] = [0, 0, 0, 0]

        # --- resolve_movement, reimplement.py:368-371,393-399
        old = [row[1 + 6 * i] for i in range(P)]
        new, new_or = [], []
        for i in range(P):
            a, pos, ori = acts[i], row[1 + 6 * i], row[1 + 6 * i + 1]
            if a == INTERACT:
                new.append(pos)
                new_or.append(ori)
            else:
                cand = pos + self._delta(a)
                new.append(cand if T[cand] == AIR else pos)
                new_or.append(ori if a == STAY else a)
        # _handle_collisions, :356-366: one colliding pair freezes every position
        blocked = False
        for i in range(P):
            for j in range(i + 1, P):
                if new[i] == new[j] or (new[i] == old[j] and old[i] == new[j]):
                    blocked = True
        for i in range(P):
            if not blocked:
                row[1 + 6 * i] = new[i]
            row[1 + 6 * i + 1] = new_or[i]

        # --- step_environment_effects, reimplement.py:373-379
        row[0] += 1
        for pos in range(S):
            o = c0 + 4 * pos
            if row[o] == SOUP and self._cooking(row[o + 1], row[o + 2], row[o + 3]):
                row[o + 3] += 1
        return rew

    # ------------------------------------------------------------------ observation
    def encode_row(self, row):
        """lossless_state_encoding (reimplement.py:173-259) followed by the adapter's
        reshape/transpose to (W, H, C) (envs/overcooked2_env.py:322-325); int8 [P,W,H,C]."""
        P, S, C, T, c0 = self.P, self.S, self.C, self.terrain, self.cell0
        sh = 5 * P
        shared = self.base.copy()
        for pos in range(S):
            o = c0 + 4 * pos
            name = row[o]
            if name == NONE:
                continue
            if name == SOUP:
                if T[pos] == POT:
                    shared[pos, sh + 5] = row[o + 1]
                    shared[pos, sh + 6] = max(int(row[o + 3]), 0)
                else:
                    shared[pos, sh + 7] = 1
            elif name == DISH:
                shared[pos, sh + 8] = 1
            elif name == ONION:
                shared[pos, sh + 9] = 1
        out = np.empty((P, self.W, self.H, C), dtype=np.int8)
        for viewer in range(P):
            plane = shared.copy()
            other = 1
            for i in range(P):
                pl = 1 + 6 * i
                pos, ori, held = row[pl], row[pl + 1], row[pl + 2]
                if i == viewer:
                    plane[pos, 0] = 1
                    plane[pos, P + ori] = 1
                else:
                    plane[pos, other] = 1
                    plane[pos, P + 4 * other + ori] = 1
                    other += 1
                if held == SOUP:
                    plane[pos, sh + 7] = 1
                elif held == DISH:
                    plane[pos, sh + 8] = 1
                elif held == ONION:
                    plane[pos, sh + 9] = 1
            out[viewer] = plane.reshape(self.H, self.W, C).transpose(1, 0, 2)
        return out

    # ------------------------------------------------------------------ batched API
    def observe(self):
        obs = np.empty((self.P, self.N, self.W, self.H, self.C), dtype=np.int8)
        for n in range(self.N):
            obs[:, n] = self.encode_row(self.state[n].tolist())
        return obs

    def step(self, actions, with_obs=True):
        """actions int [P,N] -> (obs int8 [P,N,W,H,C] | None, reward int32 [P,N], done int32 [N]).

        Episode end as in SimplifiedOvercooked.n_step (envs/overcooked2_env.py:334-339:
        done = timestep >= horizon, reward = sum over players for every player) and
        SyncVectorEnv.n_step (pantheonrl_extension/vectorenv.py:369-370: reset at once,
        return the post-reset observation)."""
        actions = np.asarray(actions).reshape(self.P, self.N)
        rew = np.zeros((self.P, self.N), dtype=np.int32)
        done = np.zeros((self.N,), dtype=np.int32)
        obs = np.empty((self.P, self.N, self.W, self.H, self.C), dtype=np.int8) if with_obs else None
        start = self.start_row()
        for n in range(self.N):
            row = self.state[n].tolist()
            acts = [int(a) if 0 <= int(a) <= 5 else STAY for a in actions[:, n]]
            r = self.step_row(row, acts)
            rew[:, n] = sum(r)
            if row[0] >= self.horizon:
                done[n] = 1
                row = list(start)
            self.state[n] = row
            if with_obs:
                obs[:, n] = self.encode_row(row)
        return obs, rew, done


# ---------------------------------------------------------------------- action RNG
_PHILOX_M0, _PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85
_M32 = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11); returns the 4 output words."""
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _M32, p1 & _M32, ((p0 >> 32) ^ c3 ^ k1) & _M32, p0 & _M32
        k0 = (k0 + _PHILOX_W0) & _M32
        k1 = (k1 + _PHILOX_W1) & _M32
    return c0, c1, c2, c3


def random_action(seed: int, world: int, step: int, player: int, num_players: int, num_actions: int = 6) -> int:
    """The device action stream of ocb_rollout_random / bb_rollout_random.

    One Philox block covers ``8 // P_pad`` consecutive steps (P_pad = 2 for P <= 2 else 4):
    counter = (world, block_lo, block_hi, 0), key = (seed_lo, seed_hi); player p of
    sub-step s uses the 16-bit slice number ``s*P_pad + p`` and maps it to
    ``(slice * num_actions) >> 16``."""
    p_pad = 2 if num_players <= 2 else 4
    steps_per_block = 8 // p_pad
    block, sub = divmod(step, steps_per_block)
    out = philox4x32_10(world & _M32, block & _M32, (block >> 32) & _M32, 0, seed & _M32, (seed >> 32) & _M32)
    h = sub * p_pad + player
    word = out[h >> 1]
    half = (word >> 16) if (h & 1) else (word & 0xFFFF)
    return (half * num_actions) >> 16


def random_actions(seed: int, world0: int, num_worlds: int, step0: int, num_steps: int, num_players: int,
                   num_actions: int = 6):
    """uint8 [K, P, N] block of the device action stream."""
    out = np.empty((num_steps, num_players, num_worlds), dtype=np.uint8)
    for k in range(num_steps):
        for n in range(num_worlds):
            for p in range(num_players):
                out[k, p, n] = random_action(seed, world0 + n, step0 + k, p, num_players, num_actions)
    return out
