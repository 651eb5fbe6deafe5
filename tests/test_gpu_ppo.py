"""GPU tests of the PPO minibatch path (ocb_minibatch_gather / ocb_policy_evaluate / ocb_ppo_loss) through the C ABI.

Parity targets: the goldens produced by the reference's own feed_forward_generator, evaluate_actions and
R_MAPPO.ppo_update (tests/golden/ppo.npz), and the numpy oracle (oracle/ppo_oracle.py, itself pinned to those goldens)
on seeded inputs.  Index / byte work is bit-exact; floating point within the stated tolerances (the north star's
1e-3 relative on anything downstream of the logits; 2e-5 on the loss arithmetic given identical inputs)."""
import os

import numpy as np
import pytest
import torch

from diverse_conventions_b200 import layouts, ppo
from diverse_conventions_b200.overcooked_env import B200Overcooked
from diverse_conventions_b200.policy import FusedPolicy, PolicyNet, log_softmax_sample
from diverse_conventions_b200.rollout import PolicyRollout
from oracle import ppo_oracle as po
from test_gpu_policy import golden_nets

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ppo.npz"))
CASES = ["default", "mse_unclipped", "no_valuenorm", "inactive", "inactive_nomask", "small_delta"]
T, N, P, NMB = int(G["T"]), int(G["N"]), int(G["P"]), int(G["num_mini_batch"])
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


class GoldenBuffer:
    """the golden rollout buffer in the seat-major layout of RolloutBuffer"""

    def __init__(self, name="default"):
        self.T, self.N, self.P = T, N, P
        obs = G["obs_seat_major"]
        self.obs = torch.zeros((T + 1,) + obs.shape[1:], dtype=torch.int8, device="cuda")
        self.obs[:T] = cu(obs)
        self.actions = cu(G[name + "_actions"])
        self.action_log_probs = cu(G[name + "_old_logp"])
        self.value_preds = torch.zeros((T + 1, P, N), dtype=torch.float32, device="cuda")
        self.value_preds[:T] = cu(G[name + "_value_preds"])
        self.returns = torch.zeros((T + 1, P, N), dtype=torch.float32, device="cuda")
        self.returns[:T] = cu(G[name + "_returns"])
        self.advantages = cu(G[name + "_adv"])
        self.active = cu(G[name + "_active"])


def test_sampler_draws_the_reference_permutation():
    gen = torch.Generator().manual_seed(int(G["seed"]))
    rows = ppo.minibatch_rows(T, N, P, NMB, gen)
    for i in range(NMB):
        assert np.array_equal(rows[i].numpy(), po.flat_to_rows(G["mb%d_flat_index" % i], N, P))


@pytest.mark.parametrize("obs_dtype", [torch.int8, torch.float32])
def test_gather_matches_the_generators_minibatch(obs_dtype):
    buf = GoldenBuffer()
    rows = cu(po.flat_to_rows(G["mb0_flat_index"], N, P))
    mb = ppo.gather_minibatch(buf, rows, buf.advantages, buf.returns, obs_dtype)
    torch.cuda.synchronize()
    assert mb["obs_batch"].dtype == obs_dtype and mb["share_obs_batch"] is mb["obs_batch"]
    assert np.array_equal(mb["obs_batch"].cpu().numpy().astype(np.int8), G["mb0_obs_batch"])
    assert np.array_equal(mb["actions_batch"].cpu().numpy(), G["mb0_actions_batch"].astype(np.int32))
    for k in ("value_preds_batch", "return_batch", "old_action_log_probs_batch", "adv_targ"):
        assert np.array_equal(mb[k].cpu().numpy(), G["mb0_" + k]), k


@pytest.mark.parametrize("W,H,R,B", [(5, 5, 50_000, 33_333),   # 500-byte rows: 4-byte but not 16-byte multiples
                                     (9, 5, 30_001, 20_002),   # 900 bytes (asymmetric_advantages), B % 4 == 2
                                     (5, 4, 40_000, 10_003),   # 400 bytes: 16-byte multiples
                                     (16, 16, 3_001, 2_047),   # 5,120 bytes: the largest grid the env accepts
                                     (3, 3, 1_000, 3)])        # 180 bytes, fewer rows than one group
def test_gather_large_ragged_against_torch_indexing(W, H, R, B):
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.randint(-3, 21, (R, W, H, 20), dtype=torch.int8, device="cuda", generator=g)

    class Buf:
        pass
    b = Buf()
    b.obs, b.actions = obs, torch.randint(0, 6, (R,), dtype=torch.int32, device="cuda", generator=g)
    b.value_preds, b.action_log_probs = torch.randn(R, device="cuda", generator=g), torch.randn(R, device="cuda", generator=g)
    ret, adv = torch.randn(R, device="cuda", generator=g), torch.randn(R, device="cuda", generator=g)
    rows = torch.randint(0, R, (B,), dtype=torch.int32, device="cuda", generator=g)
    rows[0], rows[-1] = R - 1, 0   # first and last row of the source buffer
    for dt in (torch.int8, torch.float32):
        mb = ppo.gather_minibatch(b, rows, adv, ret, dt)
        torch.cuda.synchronize()
        idx = rows.long()
        assert torch.equal(mb["obs_batch"], obs[idx].to(dt))
        assert torch.equal(mb["actions_batch"][:, 0], b.actions[idx]) and torch.equal(mb["adv_targ"][:, 0], adv[idx])
        assert torch.equal(mb["return_batch"][:, 0], ret[idx]) and torch.equal(mb["value_preds_batch"][:, 0], b.value_preds[idx])


def test_evaluate_actions_matches_the_reference_networks(golden_dir):
    g, lp, nets = golden_nets(golden_dir, "simple")
    pol = FusedPolicy(lp, 64, 1)
    pol.set_weights(0, nets["actor"], nets["critic"])
    buf = GoldenBuffer()
    for i in range(NMB):
        k = "default_mb%d_" % i
        rows = cu(po.flat_to_rows(G["mb%d_flat_index" % i], N, P))
        ev = pol.evaluate(buf.obs, buf.actions, rows, want_logits=True)
        torch.cuda.synchronize()
        assert np.allclose(ev["logits"].cpu().numpy(), G[k + "logits"], rtol=1e-3, atol=2e-6)
        assert np.allclose(ev["logp"].cpu().numpy(), G[k + "logp"], rtol=1e-3, atol=2e-6)
        assert np.allclose(ev["entropy"].cpu().numpy(), G[k + "entropy"], rtol=1e-3, atol=2e-6)
        assert np.allclose(ev["values"].cpu().numpy(), G[k + "values"], rtol=1e-3, atol=2e-5)
        # in place == on the materialised minibatch (same kernel, same tiles): bitwise
        mb = ppo.gather_minibatch(buf, rows, buf.advantages, buf.returns)
        ev2 = pol.evaluate(mb["obs_batch"], mb["actions_batch"].reshape(-1).contiguous(), None, want_logits=True)
        torch.cuda.synchronize()
        for key in ("logits", "logp", "entropy", "values"):
            assert torch.equal(ev[key], ev2[key]), key


@pytest.mark.parametrize("hidden,layout", [(64, "simple"), (64, "unident_s"), (512, "simple"), (512, "random1")])
def test_evaluate_actions_sharp_distributions_ragged(hidden, layout):
    """head gain 1 (logits far from uniform, so log-prob and entropy are exercised), a minibatch that is not a
    multiple of the 128-row tile, rows with repeats, against the fp32 torch forward"""
    lp = layouts.load_layout(layout, 400)
    actor = PolicyNet("actor", lp.width, lp.height, lp.channels, hidden).init_like_reference(3, gain=1.0)
    critic = PolicyNet("critic", lp.width, lp.height, lp.channels, hidden).init_like_reference(4)
    for net in (actor, critic):
        net.fc2_b.uniform_(-0.1, 0.1), net.head_b.uniform_(-0.2, 0.2)
    actor.head_w.mul_(6.0)
    pol = FusedPolicy(lp, hidden, 1)
    pol.set_weights(0, actor, critic)
    env = B200Overcooked(layout, 300, 0, horizon=400, seed=2)
    ro = env.rollout_random(40)
    obs, acts = ro["obs"], ro["actions"]  # [K,P,N,W,H,C], [K,P,N] uint8
    R = obs.shape[0] * obs.shape[1] * obs.shape[2]
    actions_src = acts.reshape(-1).to(torch.int32).contiguous()
    g = torch.Generator(device="cuda").manual_seed(1)
    rows = torch.randint(0, R, (1000,), dtype=torch.int32, device="cuda", generator=g)
    ev = pol.evaluate(obs, actions_src, rows, want_logits=True)
    torch.cuda.synchronize()
    x = obs.reshape(R, lp.width, lp.height, lp.channels)[rows.long()].cpu()
    ref_logits, ref_values = actor.forward(x), critic.forward(x)[:, 0]
    a = actions_src[rows.long()].cpu()
    ref_logp = log_softmax_sample(ref_logits, a)
    ref_ent = torch.distributions.Categorical(logits=ref_logits).entropy()
    scale = float(ref_logits.abs().max())
    assert float((ev["logits"].cpu() - ref_logits).abs().max()) < 1e-3 * scale
    assert float((ev["logp"].cpu() - ref_logp).abs().max()) < 2e-3 * scale
    assert float((ev["entropy"].cpu() - ref_ent).abs().max()) < 2e-3
    assert float(ref_ent.min()) < 1.6  # the distributions are not uniform
    assert float((ev["values"].cpu() - ref_values).abs().max()) < 1e-3 * float(ref_values.abs().max())
    # logp / entropy are exactly consistent with the kernel's own logits
    lp2, ent2 = po.evaluate_head(ev["logits"].cpu().numpy(), a.numpy())
    assert np.allclose(ev["logp"].cpu().numpy(), lp2, atol=5e-6) and np.allclose(ev["entropy"].cpu().numpy(), ent2, atol=5e-6)


def case_kwargs(name):
    c = G[name + "_cfg"]
    return dict(clip_param=float(c[0]), huber_delta=float(c[1]), use_clipped_value_loss=bool(c[2]), use_huber_loss=bool(c[3]),
                use_value_active_masks=bool(c[5]), use_policy_active_masks=bool(c[6])), bool(c[4])


@pytest.mark.parametrize("name", CASES)
def test_loss_and_gradients_match_the_reference_update(name):
    kw, use_vn = case_kwargs(name)
    buf = GoldenBuffer(name)
    vn = None
    if use_vn:
        vn = ppo.ValueNormState("cuda")
        vn.state.copy_(cu(G[name + "_vn_state0"]))
    inactive = name.startswith("inactive")
    for i in range(NMB):
        k = "%s_mb%d_" % (name, i)
        rows = cu(po.flat_to_rows(G["mb%d_flat_index" % i], N, P))
        out = ppo.ppo_loss(rows, cu(G[k + "logp"]), cu(G[k + "entropy"]), cu(G[k + "values"]), buf.action_log_probs,
                           buf.advantages, buf.value_preds, buf.returns, buf.active if inactive else None, vn, **kw)
        torch.cuda.synchronize()
        pl, vl, de, rm = G[k + "losses"]
        np.testing.assert_allclose(out["policy_loss"].item(), pl, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(out["value_loss"].item(), vl, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(out["dist_entropy"].item(), de, rtol=2e-6)
        np.testing.assert_allclose(out["ratio_mean"].item(), rm, rtol=2e-6)
        np.testing.assert_allclose(out["imp_weights"].cpu().numpy(), G[k + "imp_weights"], rtol=2e-6)
        np.testing.assert_allclose(out["dlogp"].cpu().numpy(), G[k + "dlogp"], rtol=2e-5, atol=1e-9)
        np.testing.assert_allclose(out["dvalues"].cpu().numpy(), G[k + "dvalues"], rtol=2e-5, atol=1e-9)
        if use_vn:
            np.testing.assert_allclose(vn.state.cpu().numpy(), G[k + "vn_state"], rtol=1e-6)


def test_full_size_minibatch_pass_over_a_collected_rollout():
    """config-4 shape: 8,192 worlds, T=25 self-play rollout, returns on the device, then two minibatches evaluated in
    place; every output is checked against the numpy oracle fed with the device's own intermediate tensors, and the
    two minibatches together cover every sample exactly once"""
    lp = layouts.load_layout("simple", 400)
    pol = FusedPolicy(lp, 64, 1)
    pol.set_weights(0, PolicyNet("actor", 5, 4, 20, 64).init_like_reference(1, gain=1.0),
                    PolicyNet("critic", 5, 4, 20, 64).init_like_reference(2))
    Nw, Tt = 8192, 25
    env = B200Overcooked("simple", Nw, 0, horizon=400, seed=3)
    ro = PolicyRollout(env, pol, Tt, seed=5)
    buf = ro.collect()
    buf.compute_returns()
    vn = ppo.ValueNormState("cuda")
    mbs = ppo.PPOMinibatches(buf, pol, num_mini_batch=2, value_norm=vn, generator=torch.Generator().manual_seed(0))
    seen = []
    state = np.zeros(3, dtype=np.float32)
    for mb in mbs:
        torch.cuda.synchronize()
        rows = mb["rows"].long()
        seen.append(rows.cpu().numpy())
        flat = lambda t: t[:Tt].reshape(-1)[rows].cpu().numpy()
        # the rollout sampled these actions from the same weights: new log-probs equal the stored ones
        assert np.allclose(mb["logp"].cpu().numpy(), flat(buf.action_log_probs), atol=1e-5)
        assert np.allclose(mb["values"].cpu().numpy(), flat(buf.value_preds), atol=1e-5)
        ref = po.ppo_loss(mb["logp"].cpu().numpy(), mb["entropy"].cpu().numpy(), mb["values"].cpu().numpy(),
                          flat(buf.action_log_probs), flat(buf.advantages), flat(buf.value_preds), flat(buf.returns), None, state)
        state = ref["vn_state"]
        for key in ("policy_loss", "value_loss", "dist_entropy", "ratio_mean"):
            np.testing.assert_allclose(mb[key].item(), ref[key], rtol=1e-5, atol=1e-7, err_msg=key)
        np.testing.assert_allclose(mb["dlogp"].cpu().numpy(), ref["dlogp"], rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(mb["dvalues"].cpu().numpy(), ref["dvalues"], rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(vn.state.cpu().numpy(), state, rtol=1e-6)
    allrows = np.concatenate(seen)
    assert allrows.size == Tt * 2 * Nw and np.array_equal(np.sort(allrows), np.arange(Tt * 2 * Nw))
