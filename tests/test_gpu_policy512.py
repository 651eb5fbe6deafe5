"""GPU tests of the hidden_size-512 policy forward (csrc/policy512.cuh: conv512_kernel + gemm512_kernel)
through the C ABI, against a plain PyTorch fp32 forward of the same networks (PolicyNet.forward, which
tests/test_policy_reference_parity.py pins to the reference's R_Actor / R_Critic at both hidden sizes).
Tolerance: the north star allows 1e-3 relative; the hi/lo bf16 operand split is asserted at 2e-4."""
import numpy as np
import pytest
import torch

from diverse_conventions_b200 import layouts
from diverse_conventions_b200.overcooked_env import B200Overcooked
from diverse_conventions_b200.policy import FusedPolicy, PolicyNet, log_softmax_sample
from diverse_conventions_b200.rollout import PolicyRollout
from oracle.c_oracle import COracle

pytestmark = pytest.mark.gpu

REL_TOL = 2e-4


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def make_nets(lp, n, seed0=31, gain=1.0):
    actors = [PolicyNet("actor", lp.width, lp.height, lp.channels, 512).init_like_reference(seed0 + i, gain=gain) for i in range(n)]
    critics = [PolicyNet("critic", lp.width, lp.height, lp.channels, 512).init_like_reference(seed0 + 100 + i) for i in range(n)]
    for net in actors + critics:
        for b in (net.conv_b, net.fc1_b, net.fc2_b, net.head_b):
            b.uniform_(-0.1, 0.1)
    return actors, critics


def play_obs(lp, N, steps=50, seed=1):
    orc = COracle(lp, N)
    rng = np.random.default_rng(seed)
    for _ in range(steps):
        o, _, _ = orc.step(rng.choice(6, size=(2, N), p=[.15, .15, .15, .15, .05, .35]))
    return torch.from_numpy(o.reshape(2 * N, lp.width, lp.height, lp.channels).copy()).cuda()


@pytest.mark.parametrize("layout,N", [("simple", 64), ("simple", 333), ("random1", 200), ("unident_s", 150)])
def test_hidden512_matches_torch_fp32(layout, N):
    lp = layouts.load_layout(layout, 400)
    actors, critics = make_nets(lp, 1)
    pol = FusedPolicy(lp, 512, 1)
    pol.set_weights(0, actors[0], critics[0])
    obs = play_obs(lp, N)
    out = pol.forward(obs, deterministic=True, want_logits=True)
    torch.cuda.synchronize()
    ref_l, ref_v = actors[0].forward(obs.cpu()), critics[0].forward(obs.cpu())[:, 0]
    assert rel_err(out["logits"].cpu(), ref_l) < REL_TOL
    assert rel_err(out["values"].cpu(), ref_v) < REL_TOL
    lg = out["logits"].cpu()
    assert torch.equal(out["actions"].cpu().long(), lg.argmax(-1))
    assert torch.allclose(out["logp"].cpu(), log_softmax_sample(lg, out["actions"].cpu()), atol=1e-6)
    # single-network entry points
    a = pol.act(obs, deterministic=True, want_logits=True)
    v = pol.value(obs)
    assert torch.equal(a["logits"], out["logits"]) and torch.equal(v, out["values"])


def test_hidden512_many_tiles_and_policy_selection():
    """more tiles than CTAs (every CTA runs several units back to back) and a per-tile policy table"""
    lp = layouts.load_layout("simple", 400)
    n_pol = 3
    actors, critics = make_nets(lp, n_pol)
    pol = FusedPolicy(lp, 512, n_pol)
    for i in range(n_pol):
        pol.set_weights(i, actors[i], critics[i])
    N = 148 * 64 + 77
    obs = play_obs(lp, N, steps=30)
    M = 2 * N
    tiles = (M + 127) // 128
    tp = torch.tensor([(t // 3) % n_pol for t in range(tiles)], dtype=torch.int32, device="cuda")
    out = pol.forward(obs, tile_policy=tp, deterministic=True, want_logits=True)
    torch.cuda.synchronize()
    for t in list(range(0, tiles, 11)) + [tiles - 1]:
        sl = slice(t * 128, min(M, (t + 1) * 128))
        k = int(tp[t])
        assert rel_err(out["logits"][sl].cpu(), actors[k].forward(obs[sl].cpu())) < REL_TOL, t
        assert rel_err(out["values"][sl].cpu(), critics[k].forward(obs[sl].cpu())[:, 0]) < REL_TOL, t


def test_hidden512_rollout_matches_oracle():
    """the device-resident rollout runs unchanged on the wide networks"""
    layout, N, T, horizon = "simple", 256, 20, 15
    lp = layouts.load_layout(layout, horizon)
    actors, critics = make_nets(lp, 1, gain=2.0)
    pol = FusedPolicy(lp, 512, 1)
    pol.set_weights(0, actors[0], critics[0])
    env = B200Overcooked(layout, N, 0, horizon=horizon, seed=5)
    ro = PolicyRollout(env, pol, T, seed=7, use_graph=False)
    buf = ro.collect()
    torch.cuda.synchronize()
    # the trajectory is what the oracle produces for the actions the policy chose
    orc = COracle(lp, N)
    first = orc.observe()
    o, r, d = orc.rollout(buf.actions.cpu().numpy().astype(np.uint8))
    assert np.array_equal(buf.obs.cpu().numpy(), np.concatenate([first[None], o]))
    assert np.array_equal(buf.rewards.cpu().numpy(), r) and np.array_equal(buf.dones.cpu().numpy(), d)
    assert int(buf.dones.sum()) > 0
    for t in (0, 7, T - 1):
        rows = buf.obs[t].reshape(-1, lp.width, lp.height, lp.channels).cpu()
        assert rel_err(buf.value_preds[t].reshape(-1).cpu(), critics[0].forward(rows)[:, 0]) < REL_TOL
        lp_ref = log_softmax_sample(actors[0].forward(rows), buf.actions[t].reshape(-1).cpu())
        assert torch.allclose(buf.action_log_probs[t].reshape(-1).cpu(), lp_ref, atol=2e-4)
    # a CUDA-graph replay of the same launch sequence (the scratch is sized by now)
    ro2 = PolicyRollout(env, pol, T, seed=7, use_graph=True)
    b2 = ro2.collect()
    torch.cuda.synchronize()
    rows = b2.obs[5].reshape(-1, lp.width, lp.height, lp.channels).cpu()
    assert rel_err(b2.value_preds[5].reshape(-1).cpu(), critics[0].forward(rows)[:, 0]) < REL_TOL
