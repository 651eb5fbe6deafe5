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
import numpy as np
import pytest
import torch

from diverse_conventions_b200 import layouts
from emu.build_emu import lib as emu_lib
from oracle import ref_shim
from oracle.c_oracle import COracle
from oracle.overcooked_oracle import OvercookedOracle
from random_layouts import random_layout, run_reference
from scripted_agent import ScriptedTeam

pytestmark = pytest.mark.reference

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
    ref = run_reference(ns, path, lp, horizon, steps, rng, ScriptedTeam)
    acts = np.ascontiguousarray(ref["actions"][:, :, None])
    ref_obs, ref_rew = ref["obs"][:, :, None], np.repeat(ref["rewards"][:, None, None], P, axis=1)
    ref_done, ref_state, reset_obs = ref["dones"][:, None], ref["states"], ref["reset_obs"]

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
