"""Pins the CPU oracle (oracle/overcooked_oracle.py and oracle/ocb_oracle.c) against
the golden vectors produced by the reference's own Python env
(tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from diverse_conventions_b200 import layouts
from oracle.c_oracle import CBalanceOracle, COracle, random_actions
from oracle.overcooked_oracle import OvercookedOracle
from oracle import overcooked_oracle as pyo

ALL = layouts.builtin_layout_names()


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, "overcooked_%s.npz" % name))


def _replay(oracle, g, steps):
    P = oracle.P
    sha = hashlib.sha256()
    assert np.array_equal(oracle.observe()[:, 0], g["reset_obs"])
    for t in range(steps):
        obs, rew, done = oracle.step(g["actions"][t].reshape(P, 1))
        assert rew[0, 0] == g["rewards"][t] and np.all(rew == rew[0, 0]), t
        assert done[0] == g["dones"][t], t
        assert np.array_equal(oracle.state[0], g["states"][t]), t
        if t < g["obs_head"].shape[0]:
            assert np.array_equal(obs[:, 0], g["obs_head"][t]), t
        sha.update(np.ascontiguousarray(obs[:, 0]).tobytes())
    return sha.digest()


@pytest.mark.parametrize("name", ALL)
def test_c_oracle_replays_reference_trajectory(golden_dir, name):
    g = _load(golden_dir, name)
    orc = COracle(layouts.load_layout(name, int(g["horizon"])), 1)
    digest = _replay(orc, g, g["actions"].shape[0])
    assert digest == g["obs_sha256"].tobytes()
    assert g["dones"].sum() == 3  # 1200 steps, horizon 400


@pytest.mark.parametrize("name", ["simple", "random0", "unident_s", "simple_tomato", "multiplayer_schelling",
                                  "simple_single"])
def test_python_oracle_replays_reference_trajectory(golden_dir, name):
    g = _load(golden_dir, name)
    orc = OvercookedOracle(layouts.load_layout(name, int(g["horizon"])), 1)
    _replay(orc, g, 450)


def test_kat(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        kat = json.load(f)
    for cls in (COracle, OvercookedOracle):
        orc = cls(layouts.load_layout("simple", 400), 1)
        obs = orc.observe()
        for p in range(2):
            nz = [[int(x), int(y), int(c), int(obs[p, 0, x, y, c])] for x, y, c in np.argwhere(obs[p, 0] != 0)]
            assert nz == kat["kat1_reset_nonzero"][p]
        sha = hashlib.sha256()
        trace = []
        for a in kat["kat2_actions_p0"]:
            obs, rew, done = orc.step(np.array([[a], [kat["kat2_action_p1"]]]))
            trace.append(int(rew[0, 0]))
            sha.update(np.ascontiguousarray(obs[:, 0]).tobytes())
        assert trace == kat["kat2_rewards"] and sum(trace) == 37
        assert sha.hexdigest() == kat["kat2_obs_sha256"]
        # SURVEY.md 8c, KAT-2 hash as probed from the reference during the survey
        assert sha.hexdigest() == "d32c67e951b3399b8fc7741403a1856c53bc89692230a767fff4f9f75820fb5d"


def test_c_and_python_oracle_agree_on_random_batches():
    rng = np.random.default_rng(5)
    for name in ("simple", "random3", "mdp_test", "multiplayer_schelling"):
        lp = layouts.load_layout(name, 37)
        a, b = COracle(lp, 6), OvercookedOracle(lp, 6)
        for t in range(120):
            act = rng.choice(6, size=(lp.num_players, 6), p=[.14, .14, .14, .14, .04, .4])
            oa, ra, da = a.step(act)
            ob, rb, db = b.step(act)
            assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(da, db)
            assert np.array_equal(a.state, b.state)


def test_action_rng_c_equals_python_and_is_uniform():
    c = random_actions(1234567890123, 5, 7, 1001, 13, 2)
    p = pyo.random_actions(1234567890123, 5, 7, 1001, 13, 2)
    assert np.array_equal(c, p)
    c4 = random_actions(99, 0, 3, 7, 9, 4)
    assert np.array_equal(c4, pyo.random_actions(99, 0, 3, 7, 9, 4))
    big = random_actions(0, 0, 4096, 0, 64, 2)
    freq = np.bincount(big.ravel(), minlength=6) / big.size
    assert big.max() == 5 and np.all(np.abs(freq - 1 / 6) < 0.004)
    # Philox4x32-10 known-answer test (Random123 kat_vectors: zero counter / zero key)
    assert pyo.philox4x32_10(0, 0, 0, 0, 0, 0) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert pyo.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == (
        0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)


def test_balance_beam_oracle_matches_reference_transitions(golden_dir):
    tr = np.load(os.path.join(golden_dir, "balance_beam.npz"))["transitions"]
    assert tr.shape == (3536, 32)
    orc = CBalanceOracle(tr.shape[0], seed=3)
    pre0, pre1 = tr[:, 0:7], tr[:, 7:14]
    # inject: loc, time and history straight from the pre-step observation
    orc.state[:, 0] = pre0[:, 0] - 2
    orc.state[:, 1] = pre1[:, 0] - 2
    orc.state[:, 2] = pre0[:, 6]
    orc.state[:, 3], orc.state[:, 4] = pre0[:, 1], pre0[:, 2]
    orc.state[:, 5], orc.state[:, 6] = pre1[:, 1], pre1[:, 2]
    assert np.array_equal(orc.observe()[0], pre0) and np.array_equal(orc.observe()[1], pre1)
    obs, rew, done = orc.step(tr[:, 14:16].T)
    assert np.array_equal(rew[0].view(np.int32), tr[:, 30]) and np.array_equal(rew[0], rew[1])
    assert np.array_equal(done, tr[:, 31])
    live = done == 0
    assert np.array_equal(obs[0][live], tr[live, 16:23]) and np.array_equal(obs[1][live], tr[live, 23:30])
    # a done world comes back freshly reset: time 2, empty history, positions on the beam
    d = done == 1
    assert np.all(obs[0][d][:, 6] == 2) and np.all(obs[0][d][:, [1, 2, 4, 5]] == 0)
    assert np.all((obs[0][d][:, 0] >= 2) & (obs[0][d][:, 0] <= 6))
