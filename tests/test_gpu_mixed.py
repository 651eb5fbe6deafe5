"""GPU tests of the mixed-play collection (ocb_rollout_mixed) through the C ABI.

The buffers the collection wrote are compared bit for bit with a step-by-step reconstruction:
  * per step, the two policies' actions / log-probs / values come from the (separately tested) policy entry points
    on the oracle's observation, with the same sampling counters the collection uses;
  * which rows the partner plays comes from the numpy restatement of the mask stream and the forcing schedule
    (oracle/mixed_oracle.py, pinned to the reference's XDPlayer / MixedAgent goldens);
  * the env stream comes from the C oracle;
  * the buffer placement comes from mixed_oracle.place (pinned to the reference's diaginsert / partinsert).
Plus: recorded log-probs / values against a plain PyTorch fp32 forward, the reference's slot-L convention, the returns
over the buffer against the numpy returns oracle, CUDA-graph replay, argument checking.
"""
import ctypes

import numpy as np
import pytest
import torch

from diverse_conventions_b200 import _native, layouts
from diverse_conventions_b200.mixed import MixedPlayCollector
from diverse_conventions_b200.overcooked_env import B200Overcooked
from diverse_conventions_b200.policy import log_softmax_sample
from oracle import mixed_oracle as mo
from oracle import returns_oracle
from oracle.c_oracle import COracle
from test_gpu_rollout import make_policies

pytestmark = pytest.mark.gpu

PARTNER_SEED_XOR = 0x9E3779B97F4A7C15  # ocb_rollout_mixed: the partner samples with seed ^ this


def reconstruct(lp, pol, L, N, seed, mix_seed, step0, state0, main=0, partner=1):
    """step-by-step reconstruction -> per-step streams [2L, ...] (seat-major) and the oracle"""
    P = 2
    orc = COracle(lp, N)
    orc.state[:] = state0
    obs = orc.observe()
    use = mo.use_partner(L, P, N, mix_seed, step0)
    tiles = (P * N + 127) // 128
    tm = torch.full((tiles,), main, dtype=torch.int32, device="cuda")
    tp = torch.full((tiles,), partner, dtype=torch.int32, device="cuda")
    st = {k: [] for k in ("obs", "actions", "logp", "values", "rewards", "dones", "played")}
    for s in range(2 * L):
        d_obs = torch.from_numpy(obs).cuda().contiguous()
        fm = pol.forward(d_obs, tile_policy=tm, seed=seed, offset=step0 + s)
        fp = pol.act(d_obs, tile_policy=tp, seed=(seed ^ PARTNER_SEED_XOR) & (2**64 - 1), offset=step0 + s)
        a_main = fm["actions"].cpu().numpy().reshape(P, N)
        a_part = fp["actions"].cpu().numpy().reshape(P, N)
        played = np.where(use[s], a_part, a_main).astype(np.int32)
        st["obs"].append(obs)
        st["actions"].append(a_main)
        st["logp"].append(fm["logp"].cpu().numpy().reshape(P, N))
        st["values"].append(fm["values"].cpu().numpy().reshape(P, N))
        st["played"].append(played)
        obs, rew, done = orc.step(played)
        st["rewards"].append(rew)
        st["dones"].append(done)
    return {k: np.stack(v) for k, v in st.items()}, use, orc


@pytest.mark.parametrize("layout,L,replicas", [("simple", 9, 5), ("random1", 6, 27)])
def test_mixed_collection_matches_stepwise_reconstruction(layout, L, replicas):
    horizon = 7  # episodes end inside the collection
    G, N = L - 1, replicas * (L - 1)
    lp = layouts.load_layout(layout, horizon)
    pol, actors, critics = make_policies(lp, 2)
    env = B200Overcooked(layout, N, 0, horizon=horizon, seed=5)
    env.rollout_random(3)  # not from a fresh reset: the collection continues from the env's current state
    torch.cuda.synchronize()
    state0, step0 = env.get_state().copy(), env.step_count
    seed, mix_seed = 77, 4242
    col = MixedPlayCollector(env, pol, L, 0, 1, seed=seed, mix_seed=mix_seed)
    buf = col.collect()
    torch.cuda.synchronize()
    assert env.step_count == step0 + 2 * L

    st, use, orc = reconstruct(lp, pol, L, N, seed, mix_seed, step0, state0)
    assert np.array_equal(env.get_state(), orc.state)       # same env stream to the very end
    free = ~np.tile(mo.schedule(L)[0], (1, replicas))        # [2L, N]
    assert use[:, 0][free].any() and use[:, 1][free].any() and not use[:, 0][free].all()
    assert (st["played"] != st["actions"]).any()             # the partner really played somewhere

    assert np.array_equal(buf.obs[:L].cpu().numpy(), mo.place(L, st["obs"], world_axis=1))
    assert np.array_equal(buf.actions.cpu().numpy(), mo.place(L, st["actions"], world_axis=1))
    assert np.array_equal(buf.action_log_probs.cpu().numpy(), mo.place(L, st["logp"], world_axis=1))
    assert np.array_equal(buf.value_preds[:L].cpu().numpy(), mo.place(L, st["values"], world_axis=1))
    assert np.array_equal(buf.rewards.cpu().numpy(), mo.place(L, st["rewards"], world_axis=1))
    assert np.array_equal(buf.dones.cpu().numpy(), mo.place(L, st["dones"], world_axis=0))
    assert buf.dones.sum().item() > 0

    # recorded rows are main-policy rows: log-probs / values of the recorded observations under policy 0 (fp32 torch)
    rows = buf.obs[:L].cpu().reshape(L * 2 * N, lp.width, lp.height, lp.channels)
    acts = buf.actions.cpu().reshape(-1)
    ref_lp = log_softmax_sample(actors[0].forward(rows), acts)
    assert torch.allclose(buf.action_log_probs.cpu().reshape(-1), ref_lp, atol=2e-4)
    ref_v = critics[0].forward(rows)[:, 0]
    assert float((buf.value_preds[:L].cpu().reshape(-1) - ref_v).abs().max() / ref_v.abs().max()) < 2e-4

    # slot L: what the reference leaves there (zeros) and the critic's value of it
    assert not buf.obs[L].any()
    v0 = critics[0].forward(torch.zeros((1, lp.width, lp.height, lp.channels)))[0, 0]
    assert torch.allclose(buf.value_preds[L].cpu(), v0.expand(2, N), rtol=1e-5, atol=1e-6)

    # episode scores kept on the device == the oracle's stream
    rs, ep = col.mp_scores()
    assert int(ep.sum()) >= int(st["dones"].sum())


def test_returns_over_the_mixed_buffer_follow_the_reference_mask_convention():
    L, replicas, horizon = 8, 16, 5
    N = replicas * (L - 1)
    lp = layouts.load_layout("simple", horizon)
    pol, _, _ = make_policies(lp, 2)
    env = B200Overcooked("simple", N, 0, horizon=horizon, seed=2)
    col = MixedPlayCollector(env, pol, L, 0, 1, seed=1, mix_seed=2)
    buf = col.collect()
    ret, adv = buf.compute_returns(normalize=False)
    torch.cuda.synchronize()
    # reference layout [L+1, N, P, 1]: masks stored at their own slot, masks[L] = 1
    v = buf.shared_buffer_views()
    assert v["masks"].shape == (L + 1, N, 2, 1) and bool((v["masks"][L] == 1).all())
    assert torch.equal(v["masks"][:L, :, 0, 0], (1 - buf.dones).float())
    # compute_returns reads masks[t+1] (shared_buffer.py:283-286): here that is 1 - dones[t+1], and masks[L] = 1
    dn = buf.dones.cpu().numpy()
    shifted = np.concatenate([dn[1:], np.zeros_like(dn[:1])])
    want_ret, want_adv = returns_oracle.compute_returns(buf.value_preds.cpu().numpy(), buf.rewards.cpu().numpy(), shifted)
    assert np.array_equal(ret[:L].cpu().numpy(), want_ret[:L])
    assert np.array_equal(adv.cpu().numpy(), want_adv)
    assert dn.sum() > 0 and buf.rewards.abs().sum().item() >= 0


def test_graph_replay_equals_direct_launches():
    L, replicas = 7, 20
    N = replicas * (L - 1)
    lp = layouts.load_layout("simple", 400)
    pol, _, _ = make_policies(lp, 2)
    out = []
    for use_graph in (False, True):
        env = B200Overcooked("simple", N, 0, horizon=11, seed=9)
        col = MixedPlayCollector(env, pol, L, 0, 1, seed=3, mix_seed=8, use_graph=use_graph)
        col.collect()
        b = col.collect()   # the second collection continues from the first one's final state
        torch.cuda.synchronize()
        out.append([t.clone() for t in (b.obs, b.actions, b.action_log_probs, b.value_preds, b.rewards, b.dones)])
        out[-1].append(torch.from_numpy(env.get_state()))
    for a, b in zip(*out):
        assert torch.equal(a.cpu(), b.cpu())


def test_argument_checks():
    lp = layouts.load_layout("simple", 400)
    pol, _, _ = make_policies(lp, 2)
    env = B200Overcooked("simple", 30, 0, horizon=400)
    with pytest.raises(ValueError):
        MixedPlayCollector(env, pol, 8)              # 30 is not a multiple of 7
    with pytest.raises(ValueError):
        MixedPlayCollector(env, pol, 7, 0, 2)        # the handle holds 2 weight sets
    lib = _native.lib()
    col = MixedPlayCollector(env, pol, 7)
    b = col.buf
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.ocb_rollout_mixed(env._h, pol._h, 7, 0, 1, p(b.obs), p(b.actions), None, None, None, None, 0, 0, 0,
                               p(col._scratch), 16, None)
    assert rc < 0 and b"scratch" in lib.ocb_last_error()
    rc = lib.ocb_rollout_mixed(env._h, pol._h, 8, 0, 1, p(b.obs), p(b.actions), None, None, None, None, 0, 0, 0,
                               p(col._scratch), col._scratch.numel(), None)
    assert rc < 0 and b"multiple" in lib.ocb_last_error()
    # actor-only collection (no critic, no log-probs) runs
    _native.check(lib.ocb_rollout_mixed(env._h, pol._h, 7, 0, 1, p(b.obs), p(b.actions), None, None, None, None, 0, 0, 0,
                                        p(col._scratch), col._scratch.numel(), None))
    torch.cuda.synchronize()
