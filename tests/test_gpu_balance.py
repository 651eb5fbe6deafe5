"""GPU parity tests of the Balance-Beam path (BASELINE config 2) through the C ABI:
against every reachable transition of the reference's PantheonLine
(tests/golden/balance_beam.npz) and against the C oracle on random rollouts."""
import os

import numpy as np
import pytest
import torch

from oracle.c_oracle import CBalanceOracle, random_actions

pytestmark = pytest.mark.gpu


def make_env(N, **kw):
    from diverse_conventions_b200.balance_env import B200BalanceBeam
    return B200BalanceBeam(N, 0, **kw)


def test_all_reference_transitions_by_state_injection(golden_dir):
    tr = np.load(os.path.join(golden_dir, "balance_beam.npz"))["transitions"]
    n = tr.shape[0]
    env = make_env(n, seed=3)
    st = env.get_state()
    pre0, pre1 = tr[:, 0:7], tr[:, 7:14]
    st[:, 0], st[:, 1], st[:, 2] = pre0[:, 0] - 2, pre1[:, 0] - 2, pre0[:, 6]
    st[:, 3], st[:, 4], st[:, 5], st[:, 6] = pre0[:, 1], pre0[:, 2], pre1[:, 1], pre1[:, 2]
    env.set_state(st)
    vobs = env.observe()
    assert np.array_equal(vobs[0].obs.cpu().numpy(), pre0) and np.array_equal(vobs[1].obs.cpu().numpy(), pre1)
    vobs, rew, done, _ = env.n_step(torch.from_numpy(tr[:, 14:16].T.copy()).reshape(2, n, 1))
    rew, done = rew.cpu().numpy(), done.cpu().numpy()
    assert rew.dtype == np.float32 and np.array_equal(rew[0].view(np.int32), tr[:, 30]) and np.array_equal(rew[0], rew[1])
    assert np.array_equal(done, tr[:, 31])
    live = done == 0
    o0, o1 = vobs[0].obs.cpu().numpy(), vobs[1].obs.cpu().numpy()
    assert np.array_equal(o0[live], tr[live, 16:23]) and np.array_equal(o1[live], tr[live, 23:30])
    d = done == 1
    assert np.all(o0[d][:, 6] == 2) and np.all(o0[d][:, [1, 2, 4, 5]] == 0)
    assert np.all((o0[d][:, 0] >= 2) & (o0[d][:, 0] <= 6)) and np.array_equal(o0[d][:, 0], o1[d][:, 3])


@pytest.mark.parametrize("N", [1, 33, 4097])
def test_random_rollout_matches_oracle(N):
    env = make_env(N, seed=21)
    orc = CBalanceOracle(N, 21)
    orc.state[:] = env.get_state()
    assert np.array_equal(np.stack([o.obs.cpu().numpy() for o in env.observe()]), orc.observe())
    step0 = 0
    for K in (1, 5, 40):
        out = env.rollout_random(K)
        torch.cuda.synchronize()
        acts = out["actions"].cpu().numpy()
        assert np.array_equal(acts, random_actions(21, 0, N, step0, K, 2, 4))
        for k in range(K):
            o, r, d = orc.step(acts[k])
            assert np.array_equal(out["obs"][k].cpu().numpy(), o), (K, k)
            assert np.array_equal(out["rewards"][k].cpu().numpy(), r) and np.array_equal(out["dones"][k].cpu().numpy(), d)
        step0 += K
    assert np.array_equal(env.get_state(), orc.state)


def test_adapter_contract_and_reset_distribution():
    from diverse_conventions_b200.vector_api import RandomVectorAgent
    N = 65536
    env = make_env(N, seed=5)
    assert env.observation_space.shape == (7,) and env.action_space.n == 4 and env.n_players == 2
    env.add_partner_agent(RandomVectorAgent(lambda: torch.randint(0, 4, (N, 1), device="cuda")))
    ob = env.reset()
    assert ob.obs.shape == (N, 7) and ob.obs.dtype == torch.int32 and ob.state is ob.obs
    assert ob.action_mask.shape == (N, 4) and ob.action_mask.all()
    ob2, rew, done, info = env.step(torch.randint(0, 4, (N, 1), device="cuda"))
    assert rew.shape == (N,) and rew.dtype == torch.float32 and done.shape == (N,) and len(info) == N
    seen = set(np.round(np.unique(rew.cpu().numpy()).astype(np.float64), 4).tolist())
    assert seen <= {1.0, -0.2, -0.4, -0.6, -0.8, -1.0, -1.2, -1.4, -1.6, -2.0, -3.0}, seen
    # reset positions: uniform on {0..4}^2 (the reference draws them with numpy's randint)
    env.hard_reset()
    st = env.get_state()
    counts = np.bincount(st[:, 0] * 5 + st[:, 1], minlength=25)
    assert counts.min() > 0.9 * N / 25 and counts.max() < 1.1 * N / 25
    assert np.all(st[:, 2] == 2) and np.all(st[:, 3:7] == 0)
    # episodes last at most 2 steps after a reset
    out = env.rollout_random(2, obs=False)
    assert bool(out["dones"][1].bool().logical_or(out["dones"][0].bool()).all())


def test_set_state_validation():
    from diverse_conventions_b200 import _native
    env = make_env(4)
    st = env.get_state()
    st[0, 0] = 7
    with pytest.raises(_native.NativeError):
        env.set_state(st)
