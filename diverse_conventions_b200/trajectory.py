"""Trajectory dump for the reference's browser replay (last item).

The flask UI replays a JSON trajectory ``{"ep_states": [[state, ...]], "ep_actions": [[joint_action, ...]],
"ep_rewards": [[r, ...]], "mdp_params": [{...}]}`` (overcooked_flask/static/js/demo/js/overcooked-replay.js:52-53,
written by overcooked-single.js:97-102,149-150 and saved by flask_app.py:109-136).  Each state is the dict
``dictToState`` reads (bundled overcooked_ai_js, static/js/demo/replay.js:3792-3828):

    {"players": [{"position": [x, y], "orientation": [dx, dy], "held_object": null | OBJ}, ...],
     "objects": [OBJ, ...], "order_list": null}        OBJ = {"name", "position": [x, y], "state"}

with ``state = [soup_type, num_items, cook_time]`` for soups (null otherwise), directions NORTH [0,-1], SOUTH [0,1],
EAST [1,0], WEST [-1,0] (replay.js:3561-3566) and joint actions as direction arrays, ``[0, 0]`` (stay) or
``"interact"`` (lookupActions, replay.js:3830-3842).

This is host-side glue for inspecting a few worlds, not a hot path: states come from ``ocb_get_state`` (packed
``int32 [N, L]``, include/ocb.h) — the same packing the oracle uses — and are converted here.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence

import numpy as np

from .layouts import LayoutParams

DIRECTIONS = ([0, -1], [0, 1], [1, 0], [-1, 0])  # Action order of the env: NORTH, SOUTH, EAST, WEST
OBJECT_NAMES = {1: "tomato", 2: "onion", 3: "dish", 4: "soup"}  # ObjectState names, envs/overcooked2_reimplement.py:46-58


def action_to_js(a: int):
    """env action index (0..5: N, S, E, W, STAY, INTERACT; reimplement.py:35-43) -> what lookupActions accepts"""
    a = int(a)
    if 0 <= a < 4:
        return list(DIRECTIONS[a])
    return "interact" if a == 5 else [0, 0]


def _object(name: int, onions: int, tomatoes: int, tick: int, x: int, y: int) -> Dict:
    obj = {"name": OBJECT_NAMES[int(name)], "position": [int(x), int(y)], "state": None}
    if name == 4:
        # the legacy replay draws single-type soups: type = the majority ingredient, count = all items,
        # cook_time = ticks cooked so far (cooking_tick is -1 while idle)
        kind = "onion" if onions >= tomatoes else "tomato"
        obj["state"] = [kind, int(onions + tomatoes), int(max(tick, 0))]
    return obj


def state_to_dict(lp: LayoutParams, packed: Sequence[int]) -> Dict:
    """one world's packed state (ocb_get_state row: timestep, per player pos/orient/held x4, per cell obj x4) -> the
    dict ``dictToState`` reads"""
    s = np.asarray(packed, dtype=np.int64)
    P, S, W = lp.num_players, lp.size, lp.width
    if s.shape != (1 + 6 * P + 4 * S,):
        raise ValueError("packed state has %s ints, expected %d" % (s.shape, 1 + 6 * P + 4 * S))
    players = []
    for i in range(P):
        pos, orient, hn, ho, ht, hk = s[1 + 6 * i: 7 + 6 * i]
        x, y = int(pos % W), int(pos // W)
        players.append({"position": [x, y], "orientation": list(DIRECTIONS[int(orient)]),
                        "held_object": _object(hn, ho, ht, hk, x, y) if hn else None})
    objects = []
    base = 1 + 6 * P
    for c in range(S):
        n, o, t, k = s[base + 4 * c: base + 4 * c + 4]
        if n:
            objects.append(_object(n, o, t, k, c % W, c // W))
    return {"players": players, "objects": objects, "order_list": None}


def terrain_rows(lp: LayoutParams) -> List[str]:
    """the ``start_grid`` argument of OvercookedTrajectoryReplay (rows of ' XPODST' with the player digits)"""
    chars = " PXODST"  # terrain codes 0..6 (envs/overcooked2_env.py:152)
    rows = []
    for y in range(lp.height):
        row = [chars[int(lp.terrain[y * lp.width + x])] for x in range(lp.width)]
        for i in range(lp.num_players):
            if int(lp.start_player_y[i]) == y:
                row[int(lp.start_player_x[i])] = str(i + 1)
        rows.append("".join(row))
    return rows


def build_trajectory(lp: LayoutParams, states: np.ndarray, actions: np.ndarray, rewards: Optional[np.ndarray] = None) -> Dict:
    """states int32 ``[T+1, L]`` (state before each step, plus the final one) or ``[T, L]``; actions ``[T, P]``;
    rewards ``[T]`` team reward -> the trajectory dict the replay UI loads"""
    states, actions = np.asarray(states), np.asarray(actions)
    T = actions.shape[0]
    if states.shape[0] not in (T, T + 1):
        raise ValueError("need one state per action (optionally plus the final state)")
    return {
        "ep_states": [[state_to_dict(lp, s) for s in states]],
        "ep_actions": [[[action_to_js(a) for a in row] for row in actions]],
        "ep_rewards": [[int(r) for r in (rewards if rewards is not None else np.zeros(T, dtype=np.int64))]],
        "mdp_params": [{"layout_name": (getattr(lp, "extra", None) or {}).get("layout_name"), "start_grid": terrain_rows(lp),
                        "cook_time": int(lp.recipe_times[0]), "delivery_reward": int(lp.recipe_values[0]),
                        "num_players": lp.num_players, "horizon": int(lp.horizon)}],
    }


class TrajectoryRecorder:
    """records one world of a ``B200Overcooked`` env step by step (``get_state`` synchronises: for demos / debugging)

        rec = TrajectoryRecorder(env, world=0)
        obs, rew, done, info = env.n_step(actions); rec.after_step(actions, rew)
        rec.save("traj.json")"""

    def __init__(self, env, world: int = 0):
        self.env, self.world = env, world
        self.states = [env.get_state()[world].copy()]
        self.actions, self.rewards = [], []

    def after_step(self, actions, rewards=None):
        a = np.asarray(actions.detach().cpu() if hasattr(actions, "detach") else actions).reshape(self.env.num_players, -1)
        self.actions.append(a[:, self.world].astype(np.int64))
        if rewards is not None:
            r = np.asarray(rewards.detach().cpu() if hasattr(rewards, "detach") else rewards)
            self.rewards.append(int(r.reshape(self.env.num_players, -1)[0, self.world]))
        self.states.append(self.env.get_state()[self.world].copy())

    def trajectory(self) -> Dict:
        return build_trajectory(self.env.layout, np.stack(self.states), np.stack(self.actions),
                                np.asarray(self.rewards) if self.rewards else None)

    def save(self, path: str) -> None:
        with open(path, "w") as f:
            json.dump(self.trajectory(), f)
