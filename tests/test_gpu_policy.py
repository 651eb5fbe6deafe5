"""GPU tests of the fused tensor-core policy forward (csrc/policy_kernels.cu) through the C ABI.

Numerics: compared with a plain PyTorch fp32 forward of the same network
(diverse_conventions_b200/policy.py: PolicyNet.forward, itself identical to the reference's
R_Actor / R_Critic on the committed golden vectors).  Tolerance of the north star: logits
within 1e-3 relative; the hi/lo bf16 operand split gets ~1e-5, the tests assert 2e-4."""
import os

import numpy as np
import pytest
import torch

from diverse_conventions_b200 import layouts
from diverse_conventions_b200.policy import FusedPolicy, PolicyNet, log_softmax_sample
from oracle.c_oracle import COracle

pytestmark = pytest.mark.gpu

REL_TOL = 2e-4  # max |err| / max |reference| (north star allows 1e-3)


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def golden_nets(golden_dir, layout):
    g = np.load(os.path.join(golden_dir, "policy_%s_h64.npz" % layout))
    lp = layouts.load_layout(layout, 400)
    nets = {}
    for kind in ("actor", "critic"):
        net = PolicyNet(kind, lp.width, lp.height, lp.channels, 64)
        net.load_state_dict({k[len(kind) + 1:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(kind + ".")})
        nets[kind] = net
    return g, lp, nets


@pytest.mark.parametrize("layout", ["simple", "random1"])
def test_matches_reference_networks_on_golden_vectors(golden_dir, layout):
    g, lp, nets = golden_nets(golden_dir, layout)
    pol = FusedPolicy(lp, 64, 1)
    pol.set_weights(0, nets["actor"], nets["critic"])
    obs = torch.from_numpy(g["obs"]).cuda()
    out = pol.act(obs, deterministic=True, want_logits=True)
    val = pol.value(obs)
    torch.cuda.synchronize()
    ref_logits, ref_values = torch.from_numpy(g["logits"]), torch.from_numpy(g["values"])[:, 0]
    assert rel_err(out["logits"].cpu(), ref_logits) < REL_TOL
    assert rel_err(val.cpu(), ref_values) < REL_TOL
    # also element-wise: |err| <= 1e-3 * |ref| + tiny absolute floor
    assert torch.allclose(out["logits"].cpu(), ref_logits, rtol=1e-3, atol=2e-6)
    # arg-max action and its log-prob (FixedCategorical.mode / log_probs)
    lg = out["logits"].cpu()
    assert torch.equal(out["actions"].cpu().long(), lg.argmax(-1))
    assert torch.allclose(out["logp"].cpu(), log_softmax_sample(lg, out["actions"].cpu()), atol=1e-6)


def test_ragged_rows_many_policies_and_tile_selection():
    lp = layouts.load_layout("simple", 400)
    n_pol = 5
    pol = FusedPolicy(lp, 64, n_pol)
    actors = [PolicyNet("actor", 5, 4, 20, 64).init_like_reference(10 + i, gain=1.0) for i in range(n_pol)]
    critics = [PolicyNet("critic", 5, 4, 20, 64).init_like_reference(50 + i) for i in range(n_pol)]
    for i in range(n_pol):
        for net in (actors[i], critics[i]):
            net.conv_b.uniform_(-0.1, 0.1), net.fc1_b.uniform_(-0.1, 0.1), net.fc2_b.uniform_(-0.1, 0.1), net.head_b.uniform_(-0.1, 0.1)
        pol.set_weights(i, actors[i], critics[i])
    # observations from random play (objects, soups, both views)
    N = 700
    orc = COracle(lp, N)
    rng = np.random.default_rng(0)
    for _ in range(150):
        o, _, _ = orc.step(rng.choice(6, size=(2, N), p=[.15, .15, .15, .15, .05, .35]))
    M = 2 * N - 37  # not a multiple of 128
    obs = torch.from_numpy(o.reshape(2 * N, 5, 4, 20)[:M].copy()).cuda()
    tiles = (M + 127) // 128
    tile_policy = torch.tensor([(3 * t + 1) % n_pol for t in range(tiles)], dtype=torch.int32, device="cuda")
    out = pol.act(obs, tile_policy=tile_policy, deterministic=True, want_logits=True)
    val = pol.value(obs, tile_policy=tile_policy)
    torch.cuda.synchronize()
    ref_l = torch.empty((M, 6))
    ref_v = torch.empty((M,))
    for t in range(tiles):
        sl = slice(t * 128, min(M, (t + 1) * 128))
        k = int(tile_policy[t])
        ref_l[sl] = actors[k].forward(obs[sl].cpu())
        ref_v[sl] = critics[k].forward(obs[sl].cpu())[:, 0]
    assert rel_err(out["logits"].cpu(), ref_l) < REL_TOL
    assert rel_err(val.cpu(), ref_v) < REL_TOL


def test_sampling_follows_the_softmax_and_is_reproducible():
    lp = layouts.load_layout("simple", 400)
    pol = FusedPolicy(lp, 64, 1)
    actor = PolicyNet("actor", 5, 4, 20, 64).init_like_reference(4, gain=3.0)  # peaked enough to test
    pol.set_weights(0, actor, PolicyNet("critic", 5, 4, 20, 64).init_like_reference(5))
    one = torch.from_numpy(COracle(lp, 1).observe()[0]).cuda()  # reset observation of player 0
    M = 1 << 16
    obs = one.expand(M, 5, 4, 20).contiguous()
    a = pol.act(obs, seed=7, offset=3, want_logits=True)
    b = pol.act(obs, seed=7, offset=3)
    c = pol.act(obs, seed=7, offset=4)
    assert torch.equal(a["actions"], b["actions"]) and not torch.equal(a["actions"], c["actions"])
    probs = torch.softmax(a["logits"][0].cpu().double(), -1)
    freq = torch.bincount(a["actions"].cpu().long(), minlength=6).double() / M
    assert float((freq - probs).abs().max()) < 4.5 * float(torch.sqrt(probs * (1 - probs) / M).max())
    assert torch.allclose(a["logp"].cpu(), log_softmax_sample(a["logits"].cpu(), a["actions"].cpu()), atol=1e-6)


def test_unsupported_shapes_fail_loudly():
    from diverse_conventions_b200 import _native
    with pytest.raises(_native.NativeError):
        FusedPolicy(layouts.load_layout("simple", 400), hidden=128)
    with pytest.raises(_native.NativeError):
        FusedPolicy(layouts.load_layout("multiplayer_schelling", 400), hidden=64)  # 4 players
    with pytest.raises(_native.NativeError):
        FusedPolicy(layouts.load_layout("corridor", 400), hidden=64)  # 9 rows high


@pytest.mark.parametrize("layout", ["simple", "random0", "random3", "unident_s", "scenario3"])
def test_fused_forward_all_layout_sizes(layout):
    """both networks in one launch, weights resident (small grids) or streamed (large grids);
    enough rows that every CTA runs several tiles back to back"""
    lp = layouts.load_layout(layout, 400)
    n_pol = 3
    pol = FusedPolicy(lp, 64, n_pol)
    actors = [PolicyNet("actor", lp.width, lp.height, 20, 64).init_like_reference(20 + i, gain=1.0) for i in range(n_pol)]
    critics = [PolicyNet("critic", lp.width, lp.height, 20, 64).init_like_reference(70 + i) for i in range(n_pol)]
    for i in range(n_pol):
        for net in (actors[i], critics[i]):
            for b in (net.conv_b, net.fc1_b, net.fc2_b, net.head_b):
                b.uniform_(-0.1, 0.1)
        pol.set_weights(i, actors[i], critics[i])
    N = 148 * 64 * 2 + 77
    orc = COracle(lp, N)
    rng = np.random.default_rng(1)
    for _ in range(60):
        o, _, _ = orc.step(rng.choice(6, size=(2, N), p=[.15, .15, .15, .15, .05, .35]))
    M = 2 * N
    obs = torch.from_numpy(o.reshape(M, lp.width, lp.height, 20).copy()).cuda()
    tiles = (M + 127) // 128
    for tp in (None, torch.tensor([(t // 5) % n_pol for t in range(tiles)], dtype=torch.int32, device="cuda")):
        out = pol.forward(obs, tile_policy=tp, deterministic=True, want_logits=True)
        torch.cuda.synchronize()
        pick = torch.arange(0, tiles, 7)
        for t in pick.tolist() + [tiles - 1]:
            sl = slice(t * 128, min(M, (t + 1) * 128))
            k = 0 if tp is None else int(tp[t])
            assert rel_err(out["logits"][sl].cpu(), actors[k].forward(obs[sl].cpu())) < REL_TOL, (layout, t)
            assert rel_err(out["values"][sl].cpu(), critics[k].forward(obs[sl].cpu())[:, 0]) < REL_TOL, (layout, t)
        # the single-network entry points ((tile, network) units; the fused launch runs both networks of a
        # tile in one CTA and sums the head in another order) agree to fp32 round-off
        a = pol.act(obs, tile_policy=tp, deterministic=True, want_logits=True)
        v = pol.value(obs, tile_policy=tp)
        assert torch.allclose(a["logits"], out["logits"], rtol=1e-5, atol=1e-6)
        assert (a["actions"] == out["actions"]).float().mean() > 0.999
        assert torch.allclose(v, out["values"], rtol=1e-5, atol=1e-6)
    info = pol.info()
    assert info["ring_slots"] >= 2 and info["smem_bytes"] <= 227 * 1024
