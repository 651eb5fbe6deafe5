"""Live differential test against the UNMODIFIED reference on RANDOM layouts (build container only: needs
/root/reference, marker `reference`).  The committed goldens pin the 21 layout files the reference ships; here random
grids (sizes, terrain mixes, 1-4 players, per-ingredient cook times / values, shaping rewards) are written as
``.layout`` files and run through

  * the reference: get_base_layout_params (envs/overcooked2_env.py:171-291) and
    SyncVectorEnv([SimplifiedOvercooked]) (pantheonrl_extension/vectorenv.py:348-425, envs/overcooked2_env.py:294-343);
  * the layout front-end, the Python oracle, the C oracle and the CPU emulation of the CUDA kernel's per-lane code

on the same action streams (scripted cooks with noise phases, so pots fill, cook and get delivered).  Everything must
agree exactly: parsed parameters, observations, rewards, dones, packed states.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from diverse_conventions_b200 import layouts
from emu.build_emu import lib as emu_lib
from oracle import ref_shim
from oracle.c_oracle import COracle
from oracle.overcooked_oracle import OvercookedOracle
from scripted_agent import ScriptedTeam

pytestmark = pytest.mark.reference

NON_WALKABLE = "XXXXXXPODST"  # weights of the solid cell kinds


def random_layout(rng) -> dict:
    """a grid the CUDA path accepts (solid border, players on interior AIR) with at least one of every station"""
    while True:
        W, H = int(rng.integers(4, 10)), int(rng.integers(4, 7))
        g = [[NON_WALKABLE[int(rng.integers(len(NON_WALKABLE)))] for _ in range(W)] for _ in range(H)]
        air = []
        for y in range(1, H - 1):
            for x in range(1, W - 1):
                if rng.random() < 0.7:
                    g[y][x] = " "
                    air.append((x, y))
        n_players = int(rng.integers(1, 5))
        flat = "".join("".join(r) for r in g)
        if len(air) < n_players + 1 or not all(c in flat for c in "PODS"):
            continue
        for i, k in enumerate(rng.permutation(len(air))[:n_players]):
            x, y = air[int(k)]
            g[y][x] = str(i + 1)
        d = {"grid": "\n".join("".join(r) for r in g), "start_order_list": None}
        kind = int(rng.integers(3))
        if kind == 0:
            d["cook_time"], d["delivery_reward"] = int(rng.integers(1, 25)), int(rng.integers(1, 60))
        elif kind == 1:
            d["onion_time"], d["tomato_time"] = int(rng.integers(1, 9)), int(rng.integers(1, 9))
            d["onion_value"], d["tomato_value"] = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        d["rew_shaping_params"] = None if rng.random() < 0.5 else {
            "PLACEMENT_IN_POT_REW": int(rng.integers(0, 7)), "DISH_PICKUP_REWARD": int(rng.integers(0, 7)),
            "SOUP_PICKUP_REWARD": int(rng.integers(0, 9)), "DISH_DISP_DISTANCE_REW": 0, "POT_DISTANCE_REW": 0,
            "SOUP_DISTANCE_REW": 0}
        return d


def pack_ref_state(env) -> np.ndarray:
    st = env.state
    P, S = env.mdp.num_players, env.mdp.size
    row = np.zeros(1 + 6 * P + 4 * S, dtype=np.int32)
    row[0] = st.timestep

    def put(at, obj):
        if obj != 0:
            row[at:at + 4] = (obj.name, obj.num_onions, obj.num_tomatoes, obj._cooking_tick)

    for i, pl in enumerate(st.players):
        row[1 + 6 * i] = pl.position
        row[1 + 6 * i + 1] = pl.orientation
        put(1 + 6 * i + 2, pl.held_object)
    for c in range(S):
        put(1 + 6 * P + 4 * c, st.objects[c])
    return row


@pytest.mark.parametrize("seed", range(24))  # placements, pickups and deliveries occur in most seeds; walled-in stations in the rest
def test_random_layout_matches_the_live_reference(seed, tmp_path):
    ns = ref_shim.load()
    rng = np.random.default_rng(1000 + seed)
    d = random_layout(rng)
    path = str(tmp_path / ("rand%d.layout" % seed))
    with open(path, "w") as f:
        f.write(repr(d))
    horizon, steps = int(rng.integers(20, 60)), 260

    # parser
    want = ns.get_base_layout_params(path, horizon)
    lp = layouts.load_layout(path, horizon)
    got = lp.as_dict()
    for k, v in want.items():
        assert got[k] == v, k
    P = lp.num_players

    # reference run, actions from scripted cooks on the reference's own state
    venv = ns.SyncVectorEnv([lambda: ns.SimplifiedOvercooked(path, horizon=horizon)], device="cpu")
    obs = venv.n_reset()
    env = venv.envs[0]
    team = ScriptedTeam(lp, rng, noise=0.2)
    acts = np.zeros((steps, P, 1), np.uint8)
    ref_obs = np.zeros((steps, P, 1, lp.width, lp.height, lp.channels), np.int8)
    ref_rew = np.zeros((steps, P, 1), np.int32)
    ref_done = np.zeros((steps, 1), np.int32)
    ref_state = np.zeros((steps, 1 + 6 * P + 4 * lp.size), np.int32)
    reset_obs = np.stack([o.obs[0].numpy() for o in obs]).astype(np.int8)
    for t in range(steps):
        team.noise = 1.0 if (t // 40) % 3 == 2 else 0.2
        a = np.asarray(team.joint(pack_ref_state(env)), dtype=np.int64)
        obs, r, dn, _ = venv.n_step(torch.from_numpy(a).reshape(P, 1, 1))
        o = np.stack([x.obs[0].numpy() for x in obs])
        assert np.array_equal(o, o.astype(np.int8))
        acts[t, :, 0], ref_obs[t, :, 0], ref_rew[t, :, 0], ref_done[t, 0] = a, o, r.numpy()[:, 0], int(dn[0])
        ref_state[t] = pack_ref_state(env)

    # C oracle
    orc = COracle(lp, 1)
    assert np.array_equal(orc.observe()[:, 0], reset_obs)
    o, r, dn = orc.rollout(acts)
    assert np.array_equal(r, ref_rew) and np.array_equal(dn, ref_done)
    assert np.array_equal(o, ref_obs)
    assert np.array_equal(orc.state[0], ref_state[-1])

    # Python oracle, step by step with states
    py = OvercookedOracle(lp, 1)
    for t in range(steps):
        o, r, dn = py.step(acts[t])
        assert np.array_equal(o, ref_obs[t]) and np.array_equal(r, ref_rew[t]) and np.array_equal(dn, ref_done[t]), t
        assert np.array_equal(py.get_state()[0], ref_state[t]), t

    # CPU emulation of the CUDA kernel's per-lane code, two lane configurations
    for G in (1, 4):
        st = COracle(lp, 1).state.copy()
        eo = np.zeros_like(ref_obs)
        er = np.zeros_like(ref_rew)
        ed = np.zeros_like(ref_done)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        rc = emu_lib().ocemu_rollout(ctypes.byref(orc.cfg), G, p(st), 1, steps, p(np.ascontiguousarray(acts)), 0, 0, 0,
                                     p(eo), p(er), p(ed), None)
        assert rc == 0
        assert np.array_equal(er, ref_rew) and np.array_equal(ed, ref_done) and np.array_equal(eo, ref_obs)
        assert np.array_equal(st[0], ref_state[-1])
    assert ref_done.sum() == steps // horizon
