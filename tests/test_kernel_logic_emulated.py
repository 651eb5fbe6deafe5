"""CPU check of the CUDA kernel's per-lane code (csrc/oc_core.cuh compiled with g++ and
driven with the kernel's tile / lane-role / phase orchestration, tests/emu/oc_emu.cpp)
against the oracle.  The real kernel is checked on the GPU in test_gpu_*.py."""
import ctypes

import numpy as np
import pytest

from diverse_conventions_b200 import layouts
from emu.build_emu import lib as emu_lib
from oracle.c_oracle import COracle, random_actions
from scripted_agent import ScriptedTeam


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def scripted_actions(lp, N, K, rng, noise=0.25):
    team = ScriptedTeam(lp, rng, noise=noise)
    shadow = COracle(lp, N)
    acts = np.zeros((K, lp.num_players, N), np.uint8)
    for k in range(K):
        for n in range(N):
            acts[k, :, n] = team.joint(shadow.state[n])
        shadow.step(acts[k])
    return acts


@pytest.mark.parametrize("name", ["simple", "random0", "random1", "random3", "unident_s", "simple_tomato",
                                  "multiplayer_schelling", "simple_single", "corridor"])
@pytest.mark.parametrize("G", [1, 2, 4, 8, 101])  # 101: one lane per world, the transition as step_pre + step_post
def test_emulated_kernel_matches_oracle(name, G):
    rng = np.random.default_rng(hash((name, G)) % 2**32)
    lp = layouts.load_layout(name, 50)
    P, N, K = lp.num_players, 11, 130
    acts = scripted_actions(lp, N, K, rng)
    orc = COracle(lp, N)
    obs_ref, rew_ref, done_ref = orc.rollout(acts)
    st = COracle(lp, N).state.copy()
    k0, parts = 0, []
    for kk in (1, 52, K - 53):  # several launches: state is stored / reloaded in between
        obs = np.zeros((kk, P, N, lp.width, lp.height, lp.channels), np.int8)
        rew = np.zeros((kk, P, N), np.int32)
        done = np.zeros((kk, N), np.int32)
        rc = emu_lib().ocemu_rollout(ctypes.byref(orc.cfg), G, _p(st), N, kk, _p(np.ascontiguousarray(acts[k0:k0 + kk])),
                                     0, k0, 0, _p(obs), _p(rew), _p(done), None)
        assert rc == 0
        parts.append((obs, rew, done))
        k0 += kk
    assert np.array_equal(np.concatenate([p[1] for p in parts]), rew_ref)
    assert np.array_equal(np.concatenate([p[2] for p in parts]), done_ref)
    assert np.array_equal(np.concatenate([p[0] for p in parts]), obs_ref)
    assert np.array_equal(st, orc.state)
    assert rew_ref.sum() > 0 and done_ref.sum() == 2 * N


def test_emulated_rng_stream_matches_oracle_stream():
    lp = layouts.load_layout("simple", 400)
    N, K = 40, 37
    orc = COracle(lp, N)
    st = orc.state.copy()
    ao = np.zeros((K, 2, N), np.uint8)
    obs = np.zeros((K, 2, N, 5, 4, 20), np.int8)
    emu_lib().ocemu_rollout(ctypes.byref(orc.cfg), 4, _p(st), N, K, None, 77, 5, 100, _p(obs), None, None, _p(ao))
    assert np.array_equal(ao, random_actions(77, 100, N, 5, K, 2))
    o, _, _ = orc.rollout(ao)
    assert np.array_equal(o, obs)


@pytest.mark.parametrize("name", layouts.builtin_layout_names())
def test_two_half_transition_matches_oracle_on_random_play(name):
    """step_pre + step_post (what the fused rollout's env warps run) over long random play — both players facing the same
    counter or pot, pots filling and cooking, episode ends — against the C oracle: observations, rewards, final states"""
    lp = layouts.load_layout(name, 120)
    if lp.num_players != 2:
        pytest.skip("two players")
    N, K = 96, 700
    orc = COracle(lp, N)
    st = orc.state.copy()
    ao = np.zeros((K, 2, N), np.uint8)
    obs = np.zeros((K, 2, N, lp.width, lp.height, lp.channels), np.int8)
    rew = np.zeros((K, 2, N), np.int32)
    done = np.zeros((K, N), np.int32)
    assert emu_lib().ocemu_rollout(ctypes.byref(orc.cfg), 101, _p(st), N, K, None, 5, 0, 0, _p(obs), _p(rew), _p(done), _p(ao)) == 0
    o, r, d = orc.rollout(ao)
    assert np.array_equal(o, obs) and np.array_equal(r, rew) and np.array_equal(d, done)
    assert np.array_equal(st, orc.state)


def test_mixed_play_schedule_and_mask_stream_of_the_kernels_match_the_oracle():
    """csrc/mixed_schedule.h (what mix_select_kernel / mix_record_kernel execute) against oracle/mixed_oracle.py, which is
    pinned to the reference's collect_mp_episode / MixedAgent / diaginsert / partinsert goldens"""
    import ctypes

    import numpy as np

    from emu import build_emu
    from oracle import mixed_oracle as mo
    L_ = build_emu.lib()
    for L in (2, 3, 7, 10):
        G = L - 1
        forced, slot = mo.schedule(L)
        for s in range(2 * L):
            for j in range(G):
                assert bool(L_.ocemu_mix_forced(L, s, j)) == bool(forced[s, j]), (L, s, j)
            for replicas, P in ((1, 2), (3, 2)):
                N = replicas * G
                buf = np.full((2 * N * P + 16, 3), -1, dtype=np.int32)
                n = L_.ocemu_mix_items(L, s, N, P, buf.ctypes.data_as(ctypes.c_void_p), buf.shape[0])
                want = sorted((p, w, int(slot[s, w % G])) for p in range(P) for w in range(N) if forced[s, w % G])
                assert n == len(want) and sorted(map(tuple, buf[:n].tolist())) == want, (L, s, replicas)
    rng = np.random.default_rng(0)
    for _ in range(300):
        seed, row, step = int(rng.integers(0, 2**63)), int(rng.integers(0, 2**31)), int(rng.integers(0, 2**40))
        assert bool(L_.ocemu_mix_draw(seed, row, step)) == mo.mix_draw(seed, row, step)
