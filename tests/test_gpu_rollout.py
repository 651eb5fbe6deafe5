"""GPU tests of the device-resident self-play / cross-play rollout (ocb_rollout_policy).

Checks, through the C ABI:
  * the trajectory the kernels wrote (observations, rewards, dones) is bit-identical to the
    CPU oracle replaying the actions the policy kernel chose;
  * at every step the recorded log-probs / values / actions are what a plain PyTorch fp32
    forward of the same networks gives on the recorded observations (tolerance below);
  * cross-play slices run the weight sets the pair table names, and the per-pair return
    matrix equals the oracle's returns;
  * a CUDA-graph replay is equivalent to issuing the launches one by one.
"""
import numpy as np
import pytest
import torch

from diverse_conventions_b200 import layouts, sharding
from diverse_conventions_b200.overcooked_env import B200Overcooked
from diverse_conventions_b200.policy import FusedPolicy, PolicyNet, log_softmax_sample
from diverse_conventions_b200.rollout import CrossPlayEvaluator, PolicyRollout, pair_tile_policy
from oracle.c_oracle import COracle

pytestmark = pytest.mark.gpu

ATOL_LOGP = 2e-4  # log-probs are O(1); logits agree to ~1e-5 relative (north star: 1e-3)
REL_TOL = 2e-4


def make_policies(lp, n, gain=2.0, seed0=11):
    actors = [PolicyNet("actor", lp.width, lp.height, lp.channels, 64).init_like_reference(seed0 + i, gain=gain)
              for i in range(n)]
    critics = [PolicyNet("critic", lp.width, lp.height, lp.channels, 64).init_like_reference(seed0 + 500 + i)
               for i in range(n)]
    for net in actors + critics:
        for b in (net.conv_b, net.fc1_b, net.fc2_b, net.head_b):
            b.uniform_(-0.1, 0.1)
    pol = FusedPolicy(lp, 64, n)
    for i in range(n):
        pol.set_weights(i, actors[i], critics[i])
    return pol, actors, critics


def replay_through_oracle(lp, N, buf):
    orc = COracle(lp, N)
    first = orc.observe()
    o, r, d = orc.rollout(buf.actions.cpu().numpy().astype(np.uint8))
    return np.concatenate([first[None], o]), r, d, orc


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("layout,N,T", [("simple", 384, 70), ("random1", 200, 45)])
def test_selfplay_rollout_matches_oracle_and_torch(layout, N, T, fused):
    horizon = 30  # several auto-resets inside one rollout
    lp = layouts.load_layout(layout, horizon)
    pol, actors, critics = make_policies(lp, 1)
    env = B200Overcooked(layout, N, 0, horizon=horizon, seed=5)
    ro = PolicyRollout(env, pol, T, seed=123, fused=fused)
    assert ro.fused == fused
    buf = ro.collect()
    torch.cuda.synchronize()

    obs, rew, done, orc = replay_through_oracle(lp, N, buf)
    assert np.array_equal(buf.obs.cpu().numpy(), obs)
    assert np.array_equal(buf.rewards.cpu().numpy(), rew) and np.array_equal(buf.dones.cpu().numpy(), done)
    assert np.array_equal(env.get_state(), orc.state)
    assert done.sum() == N * (T // horizon)

    acts = buf.actions.cpu()
    assert int(acts.min()) >= 0 and int(acts.max()) <= 5
    rows = buf.obs.cpu().reshape(T + 1, 2 * N, lp.width, lp.height, lp.channels)
    for t in list(range(0, T, 9)) + [T - 1]:
        logits = actors[0].forward(rows[t])
        ref_lp = log_softmax_sample(logits, acts[t].reshape(-1))
        assert torch.allclose(buf.action_log_probs[t].cpu().reshape(-1), ref_lp, atol=ATOL_LOGP), t
        ref_v = critics[0].forward(rows[t])[:, 0]
        err = (buf.value_preds[t].cpu().reshape(-1) - ref_v).abs().max() / ref_v.abs().max()
        assert float(err) < REL_TOL, t
    ref_v = critics[0].forward(rows[T])[:, 0]
    assert float((buf.value_preds[T].cpu().reshape(-1) - ref_v).abs().max() / ref_v.abs().max()) < REL_TOL
    # sampled, not arg-max: a stochastic policy with gain 2 must not always pick the mode
    logits0 = actors[0].forward(rows[0])
    assert not torch.equal(acts[0].reshape(-1).long(), logits0.argmax(-1))

    # device-side score keeping == oracle episode returns
    rs, ep = env.episode_stats()
    assert int(ep.sum()) == int(done.sum())
    ret = np.zeros(N, dtype=np.int64)
    total = 0
    for t in range(T):
        ret += rew[t, 0]
        total += int(ret[done[t] != 0].sum())
        ret[done[t] != 0] = 0
    assert int(rs.sum()) == total

    # the second rollout continues from the last observation (after_update semantics)
    last = buf.obs[T].clone()
    buf2 = ro.collect()
    torch.cuda.synchronize()
    assert torch.equal(buf2.obs[0], last)
    assert ro.rollouts == 2 and env.step_count == 2 * T


@pytest.mark.parametrize("layout,N,T", [("simple", 8192, 100), ("simple", 9472 + 37, 40), ("random1", 8192, 60)])
def test_fused_rollout_at_the_benchmarked_shape_replays_through_the_oracle(layout, N, T):
    """BASELINE config 4's shape (8,192 worlds = 128 CTAs; 9,509 worlds = 148 CTAs + a second, ragged tile on one of them):
    every observation, reward and done the fused kernel wrote is what the oracle produces for the sampled actions, and a
    strided subset of rows agrees with the fp32 torch forward"""
    horizon = 23
    lp = layouts.load_layout(layout, horizon)
    pol, actors, critics = make_policies(lp, 1)
    env = B200Overcooked(layout, N, 0, horizon=horizon, seed=11)
    ro = PolicyRollout(env, pol, T, seed=5, fused=True)
    buf = ro.collect()
    torch.cuda.synchronize()
    obs, rew, done, orc = replay_through_oracle(lp, N, buf)
    assert np.array_equal(buf.rewards.cpu().numpy(), rew) and np.array_equal(buf.dones.cpu().numpy(), done)
    got = buf.obs.cpu().numpy()
    assert np.array_equal(got, obs)
    assert np.array_equal(env.get_state(), orc.state)
    assert done.sum() == N * (T // horizon) and rew.sum() > 0
    rows = torch.from_numpy(got).reshape(T + 1, 2 * N, lp.width, lp.height, lp.channels)
    sub = slice(0, 2 * N, 61)
    for t in (0, T // 2, T - 1):
        ref_lp = log_softmax_sample(actors[0].forward(rows[t, sub]), buf.actions[t].cpu().reshape(-1)[sub])
        assert torch.allclose(buf.action_log_probs[t].cpu().reshape(-1)[sub], ref_lp, atol=ATOL_LOGP), t
    ref_v = critics[0].forward(rows[T, sub])[:, 0]
    assert float((buf.value_preds[T].cpu().reshape(-1)[sub] - ref_v).abs().max() / ref_v.abs().max()) < REL_TOL


@pytest.mark.parametrize("layout,N,T", [("multiplayer_schelling", 96, 30), ("corridor", 70, 24), ("schelling", 128, 20)])
def test_rollout_on_layouts_outside_the_tensor_core_range(layout, N, T):
    """4 players / 7- and 9-row grids: the per-step rollout (generic policy kernel + the P <= 4 env kernels) writes the
    trajectory the oracle reproduces from the sampled actions, with log-probs / values of the fp32 torch networks"""
    horizon = 11
    lp = layouts.load_layout(layout, horizon)
    P = lp.num_players
    actor = PolicyNet("actor", lp.width, lp.height, lp.channels, 64).init_like_reference(21, gain=2.0)
    critic = PolicyNet("critic", lp.width, lp.height, lp.channels, 64).init_like_reference(22)
    pol = FusedPolicy(lp, 64, 1)
    pol.set_weights(0, actor, critic)
    env = B200Overcooked(layout, N, 0, horizon=horizon, seed=3)
    ro = PolicyRollout(env, pol, T, seed=17)
    assert not ro.fused or P == 2  # the persistent kernel declines; collect() falls back to per-step launches
    buf = ro.collect()
    torch.cuda.synchronize()
    assert not ro.fused
    obs, rew, done, orc = replay_through_oracle(lp, N, buf)
    assert np.array_equal(buf.obs.cpu().numpy(), obs)
    assert np.array_equal(buf.rewards.cpu().numpy(), rew) and np.array_equal(buf.dones.cpu().numpy(), done)
    assert np.array_equal(env.get_state(), orc.state) and done.sum() == N * (T // horizon)
    rows = buf.obs.cpu().reshape(T + 1, P * N, lp.width, lp.height, lp.channels)
    for t in (0, T // 2, T - 1):
        ref_lp = log_softmax_sample(actor.forward(rows[t]), buf.actions[t].cpu().reshape(-1))
        assert torch.allclose(buf.action_log_probs[t].cpu().reshape(-1), ref_lp, atol=ATOL_LOGP), t
    ref_v = critic.forward(rows[T])[:, 0]
    assert float((buf.value_preds[T].cpu().reshape(-1) - ref_v).abs().max() / ref_v.abs().max()) < REL_TOL
    v = buf.shared_buffer_views()
    assert v["obs"].shape == (T + 1, N, P, lp.width, lp.height, lp.channels) and v["masks"].shape == (T + 1, N, P, 1)


def test_shared_buffer_views_follow_the_reference_axis_order():
    lp = layouts.load_layout("simple", 400)
    pol, _, _ = make_policies(lp, 1)
    N, T = 128, 6
    env = B200Overcooked("simple", N, 0, horizon=400, seed=1)
    buf = PolicyRollout(env, pol, T).collect()
    torch.cuda.synchronize()
    v = buf.shared_buffer_views()
    assert v["obs"].shape == (T + 1, N, 2, 5, 4, 20) and v["share_obs"] is v["obs"]
    assert v["actions"].shape == (T, N, 2, 1) and v["action_log_probs"].shape == (T, N, 2, 1)
    assert v["value_preds"].shape == (T + 1, N, 2, 1) and v["rewards"].shape == (T, N, 2, 1)
    assert v["masks_next"].shape == (T, N, 2, 1)
    assert v["obs"].data_ptr() == buf.obs.data_ptr()  # a view, not a copy
    assert torch.equal(v["actions"][3, 17, 1, 0], buf.actions[3, 1, 17])
    assert torch.equal(v["obs"][2, 5, 1], buf.obs[2, 1, 5])


@pytest.mark.parametrize("fused", [False, True])
def test_deterministic_rollout_graph_replay_equals_plain_launches(fused):
    lp = layouts.load_layout("random0", 25)
    pol, _, _ = make_policies(lp, 1)
    N, T = 256, 40
    outs = []
    for use_graph in (False, True):
        env = B200Overcooked("random0", N, 0, horizon=25, seed=2)
        ro = PolicyRollout(env, pol, T, use_graph=use_graph, seed=9, fused=fused)
        a = ro.collect(deterministic=True)
        first = (a.obs.clone(), a.actions.clone(), a.value_preds.clone(), a.rewards.clone(), a.dones.clone())
        b = ro.collect(deterministic=True)
        torch.cuda.synchronize()
        outs.append((first, (b.obs.clone(), b.actions.clone(), b.value_preds.clone(), b.rewards.clone(), b.dones.clone()),
                     env.get_state()))
        env.close()
    for k in range(2):
        for x, y in zip(outs[0][k], outs[1][k]):
            assert torch.equal(x, y)
    assert np.array_equal(outs[0][2], outs[1][2])


@pytest.mark.parametrize("fused", [False, True])
def test_sampled_graph_replays_draw_fresh_actions_and_stay_exact(fused):
    lp = layouts.load_layout("simple", 400)
    pol, actors, _ = make_policies(lp, 1)
    N, T = 256, 12
    env = B200Overcooked("simple", N, 0, horizon=400, seed=4)
    ro = PolicyRollout(env, pol, T, use_graph=True, seed=77, fused=fused)
    orc = COracle(lp, N)
    prev_actions = None
    for k in range(3):
        buf = ro.collect()
        torch.cuda.synchronize()
        o, r, d = orc.rollout(buf.actions.cpu().numpy().astype(np.uint8))
        assert np.array_equal(buf.obs[1:].cpu().numpy(), o) and np.array_equal(buf.rewards.cpu().numpy(), r)
        if prev_actions is not None:
            assert not torch.equal(prev_actions[0], buf.actions[0])
        prev_actions = buf.actions.clone()
    assert np.array_equal(env.get_state(), orc.state)


@pytest.mark.parametrize("layout,N,T,horizon,index", [
    ("simple", 64 * 150 + 37, 9, 7, 0),   # more world tiles than SMs (two per CTA on some) + a ragged last tile
    ("simple", 1000, 33, 12, 1),          # second weight set of the handle
    ("unident_s", 300, 21, 10, 0),        # 9 x 5 grid: weights streamed through the ring
    ("random3", 130, 17, 400, 0),         # 8 x 5 grid, no reset inside the rollout
])
@pytest.mark.parametrize("slots", ["1", "2"])
def test_fused_rollout_is_bit_identical_to_the_per_step_launches(monkeypatch, layout, N, T, horizon, index, slots):
    """the single persistent launch (ocb_rollout_policy_fused) and the 2T+1 launches of ocb_rollout_policy fill the
    same buffers: same observations / rewards / dones, same sampled actions, same log-probs and values (bitwise),
    same final state and episode statistics, over two consecutive rollouts"""
    monkeypatch.setenv("OCB_FUSED_SLOTS", slots)  # world tiles in flight per CTA (2 falls back to 1 where the planes do not fit)
    lp = layouts.load_layout(layout, horizon)
    pol, _, _ = make_policies(lp, 2)
    res = []
    for fused in (False, True):
        env = B200Overcooked(layout, N, 0, horizon=horizon, seed=8)
        ro = PolicyRollout(env, pol, T, seed=31, fused=fused, policy_index=index)
        out = []
        for _ in range(2):
            b = ro.collect()
            torch.cuda.synchronize()
            out.append([x.clone() for x in (b.obs, b.actions, b.action_log_probs, b.value_preds, b.rewards, b.dones)])
        assert ro.fused == fused
        rs, ep = env.episode_stats()
        res.append((out, env.get_state(), rs.clone(), ep.clone(), env.step_count))
        env.close()
    names = ("obs", "actions", "logp", "values", "rewards", "dones")
    for k in range(2):
        for name, x, y in zip(names, res[0][0][k], res[1][0][k]):
            assert torch.equal(x, y), (k, name)
    assert np.array_equal(res[0][1], res[1][1])
    assert torch.equal(res[0][2], res[1][2]) and torch.equal(res[0][3], res[1][3]) and res[0][4] == res[1][4] == 2 * T


def test_fused_rollout_falls_back_when_the_layout_does_not_fit():
    """scenario3 (10 x 6): the resident observation planes leave no room for the weight ring"""
    from diverse_conventions_b200 import _native
    lp = layouts.load_layout("scenario3", 50)
    pol, _, _ = make_policies(lp, 1)
    env = B200Overcooked("scenario3", 96, 0, horizon=50, seed=1)
    with pytest.raises(_native.NativeError):
        PolicyRollout(env, pol, 4, fused=True).collect()
    ro = PolicyRollout(env, pol, 4)  # fused=None: use it when it applies
    buf = ro.collect()
    torch.cuda.synchronize()
    assert ro.fused is False
    o, r, d, _ = replay_through_oracle(lp, 96, buf)
    assert np.array_equal(buf.obs.cpu().numpy(), o) and np.array_equal(buf.rewards.cpu().numpy(), r)


def test_crossplay_slices_use_the_named_policies_and_return_matrix_matches_oracle():
    layout, horizon, wpp, n_pol = "random1", 40, 128, 3
    lp = layouts.load_layout(layout, horizon)
    pol, actors, _ = make_policies(lp, n_pol, gain=3.0)
    pairs = sharding.all_pairs(n_pol)
    table = pair_tile_policy(pairs, wpp)
    assert table.tolist() == [p[0] for p in pairs] + [p[1] for p in pairs]

    ev = CrossPlayEvaluator(layout, pol, pairs, worlds_per_pair=wpp, horizon=horizon, seed=3, chunk_steps=20,
                            use_graph=True, fused=False)  # chunked per-step launches: the trajectory is recorded below
    N = ev.env.num_envs
    # record the whole episode by chunks to replay it
    ev.env.n_reset()
    ev.env.clear_episode_stats()
    ev.rollout._primed = False
    acts, obs = [], []
    for _ in range(horizon // 20):
        b = ev.rollout.collect()
        torch.cuda.synchronize()
        acts.append(b.actions.clone())
        obs.append(b.obs[:-1].clone())
    acts, obs = torch.cat(acts).cpu(), torch.cat(obs).cpu()
    rs, ep = ev.env.episode_stats()
    rs, ep = rs.view(len(pairs), wpp).sum(1), ep.view(len(pairs), wpp).sum(1)

    # (1) environment side: oracle replay gives the same per-pair returns
    orc = COracle(lp, N)
    _, rew, done = orc.rollout(acts.numpy().astype(np.uint8), with_obs=False)
    assert done[-1].all() and done[:-1].sum() == 0
    ref_ret = rew[:, 0].sum(0).reshape(len(pairs), wpp).sum(1)
    assert np.array_equal(rs.cpu().numpy(), ref_ret) and (ep.cpu().numpy() == wpp).all()

    # (2) policy side: seat s of slice k acted with policy pairs[k][s]; the sampled actions follow
    # that policy's distribution (mean log-prob under the right actor beats every wrong actor)
    rows = obs.reshape(horizon, 2, len(pairs), wpp, lp.width, lp.height, lp.channels)
    a = acts.reshape(horizon, 2, len(pairs), wpp)
    for k in (1, 5, 6):
        for seat in (0, 1):
            o = rows[::4, seat, k].reshape(-1, lp.width, lp.height, lp.channels)
            aa = a[::4, seat, k].reshape(-1)
            scores = [float(log_softmax_sample(actors[q].forward(o), aa).mean()) for q in range(n_pol)]
            assert int(np.argmax(scores)) == pairs[k][seat], (k, seat, scores)

    # (3) run() + single-rank matrix assembly
    rs2, ep2 = ev.run()
    mean, eps = sharding.gather_pair_matrix(pairs, rs2, ep2, n_pol)
    assert mean.shape == (n_pol, n_pol) and int(eps.sum()) == N
    assert torch.isfinite(mean).all()
    ev.close()

    # (4) the whole episode as ONE persistent launch without a trajectory buffer: a fresh env draws the same streams as the
    # recorded first episode above (sampling is keyed by seed, row and the env's step counter), hence the same returns
    ev1 = CrossPlayEvaluator(layout, pol, pairs, worlds_per_pair=wpp, horizon=horizon, seed=3, chunk_steps=20)
    rs3, ep3 = ev1.run()
    torch.cuda.synchronize()
    assert ev1.fused and ev1.rollout is None
    assert torch.equal(rs3, rs) and torch.equal(ep3, ep)
    ev1.close()


@pytest.mark.parametrize("slots", ["1", "2"])
@pytest.mark.parametrize("layout,n_pol,wpp,T,horizon", [("random1", 3, 128, 26, 11), ("simple", 2, 256, 19, 400),
                                                         ("random0", 4, 128, 13, 6)])
def test_fused_crossplay_is_bit_identical_to_the_per_step_launches(monkeypatch, layout, n_pol, wpp, T, horizon, slots):
    """ocb_rollout_crossplay_fused (one persistent launch; OCB_FUSED_SLOTS forces one or two world tiles in flight per CTA)
    against ocb_rollout_policy with the same tile_policy table: same sampled actions and log-probs (bitwise), same
    observations / rewards / dones, same final state and episode statistics, over two consecutive rollouts"""
    monkeypatch.setenv("OCB_FUSED_SLOTS", slots)
    lp = layouts.load_layout(layout, horizon)
    pol, _, _ = make_policies(lp, n_pol, gain=2.0)
    pairs = sharding.all_pairs(n_pol)
    N = len(pairs) * wpp
    res = []
    for fused in (False, True):
        env = B200Overcooked(layout, N, 0, horizon=horizon, seed=4)
        ro = PolicyRollout(env, pol, T, pair_tile_policy(pairs, wpp, env.sim_device), with_critic=False, seed=77, fused=fused)
        out = []
        for _ in range(2):
            b = ro.collect()
            torch.cuda.synchronize()
            out.append([x.clone() for x in (b.obs, b.actions, b.action_log_probs, b.rewards, b.dones)])
        assert ro.fused == fused
        rs, ep = env.episode_stats()
        res.append((out, env.get_state(), rs.clone(), ep.clone(), env.step_count))
        env.close()
    for k in range(2):
        for name, x, y in zip(("obs", "actions", "logp", "rewards", "dones"), res[0][0][k], res[1][0][k]):
            assert torch.equal(x, y), (k, name)
    assert np.array_equal(res[0][1], res[1][1])
    assert torch.equal(res[0][2], res[1][2]) and torch.equal(res[0][3], res[1][3]) and res[0][4] == res[1][4] == 2 * T


def test_config5_at_full_size_replays_through_the_oracle():
    """BASELINE config 5 at its stated size — 16 x 16 policy pairs x 1,024 worlds of coordination_ring, one 400-step episode
    per world — as ONE fused launch with two tiles in flight per SM: the sampled actions are recorded, replayed through the
    CPU oracle, and every reward, every done and the per-pair returns the device accumulated must come out identical.
    (Size-independent property of the measured configuration; the policy side is pinned by the bit-identity test against the
    per-step launches at small sizes.)"""
    layout, horizon, wpp, n_pol = "random1", 400, 1024, 16
    lp = layouts.load_layout(layout, horizon)
    pol, _, _ = make_policies(lp, n_pol, gain=2.0)
    pairs = sharding.all_pairs(n_pol)
    N = len(pairs) * wpp
    env = B200Overcooked(layout, N, 0, horizon=horizon, seed=6)
    import ctypes
    from diverse_conventions_b200 import _native
    from diverse_conventions_b200.overcooked_env import _ptr
    lib = _native.lib()
    table = pair_tile_policy(pairs, wpp, env.sim_device)
    actions = torch.empty((horizon, 2, N), dtype=torch.int32, device=env.sim_device)
    rewards = torch.empty((horizon, 2, N), dtype=torch.int32, device=env.sim_device)
    dones = torch.empty((horizon, N), dtype=torch.int32, device=env.sim_device)
    env.n_reset()
    env.clear_episode_stats()
    stream = ctypes.c_void_p(torch.cuda.current_stream(env.sim_device).cuda_stream)
    _native.check(lib.ocb_rollout_crossplay_fused(env._h, pol._h, horizon, _ptr(table), None, _ptr(actions), None, _ptr(rewards),
                                                  _ptr(dones), 0, 17, stream))
    torch.cuda.synchronize()
    rs, ep = env.episode_stats()
    orc = COracle(lp, N)
    _, rew, done = orc.rollout(actions.cpu().numpy().astype(np.uint8), with_obs=False)
    assert np.array_equal(rewards.cpu().numpy(), rew) and np.array_equal(dones.cpu().numpy(), done)
    assert done[-1].all() and done[:-1].sum() == 0
    assert np.array_equal(env.get_state(), orc.state)
    ref = rew[:, 0].astype(np.int64).sum(0)
    assert np.array_equal(rs.cpu().numpy(), ref) and (ep.cpu().numpy() == 1).all()
    per_pair = ref.reshape(len(pairs), wpp).sum(1)
    assert per_pair.min() > 0  # random-init policies with gain 2 do cook on coordination_ring
    env.close()


def test_two_tiles_in_flight_at_a_large_batch_replay_through_the_oracle():
    """32,768 worlds of cramped_room (512 tiles on 148 SMs: the launch keeps two tiles in flight per CTA): every observation,
    reward and done of the self-play rollout is what the oracle produces for the sampled actions"""
    layout, N, T, horizon = "simple", 32768, 24, 11
    lp = layouts.load_layout(layout, horizon)
    pol, actors, critics = make_policies(lp, 1)
    env = B200Overcooked(layout, N, 0, horizon=horizon, seed=2)
    ro = PolicyRollout(env, pol, T, seed=9, fused=True)
    buf = ro.collect()
    torch.cuda.synchronize()
    obs, rew, done, orc = replay_through_oracle(lp, N, buf)
    assert np.array_equal(buf.rewards.cpu().numpy(), rew) and np.array_equal(buf.dones.cpu().numpy(), done)
    assert np.array_equal(buf.obs.cpu().numpy(), obs)
    assert np.array_equal(env.get_state(), orc.state)
    rows = torch.from_numpy(obs).reshape(T + 1, 2 * N, lp.width, lp.height, lp.channels)
    sub = slice(0, 2 * N, 257)
    for t in (0, T - 1):
        ref_lp = log_softmax_sample(actors[0].forward(rows[t, sub]), buf.actions[t].cpu().reshape(-1)[sub])
        assert torch.allclose(buf.action_log_probs[t].cpu().reshape(-1)[sub], ref_lp, atol=ATOL_LOGP), t
    ref_v = critics[0].forward(rows[T, sub])[:, 0]
    assert float((buf.value_preds[T].cpu().reshape(-1)[sub] - ref_v).abs().max() / ref_v.abs().max()) < REL_TOL
    env.close()


def test_actor_only_rollout_and_argument_checks():
    lp = layouts.load_layout("simple", 400)
    pol, _, _ = make_policies(lp, 2)
    env = B200Overcooked("simple", 256, 0, horizon=400, seed=1)
    ro = PolicyRollout(env, pol, 5, with_critic=False, with_logp=False)
    buf = ro.collect()
    torch.cuda.synchronize()
    assert buf.value_preds is None and buf.action_log_probs is None
    o, r, _, _ = replay_through_oracle(lp, 256, buf)
    assert np.array_equal(buf.obs.cpu().numpy(), o)
    with pytest.raises(ValueError):
        PolicyRollout(env, pol, 5, tile_policy=torch.zeros(3, dtype=torch.int32))
    with pytest.raises(ValueError):
        pair_tile_policy([(0, 1)], 100)
    other = FusedPolicy(layouts.load_layout("random1", 400), 64, 1)
    with pytest.raises(ValueError):
        PolicyRollout(env, other, 5)


def test_destroying_handles_during_a_capture_keeps_the_graph_valid():
    """handles may be garbage-collected while another rollout is being captured (cudaFree is not capturable in
    global mode): the destroy entry points switch the thread to relaxed capture mode for their frees"""
    lp = layouts.load_layout("simple", 400)
    pol, _, _ = make_policies(lp, 1)
    env = B200Overcooked("simple", 256, 0, horizon=400, seed=4)
    ro = PolicyRollout(env, pol, 6, seed=1)
    ro.prime()
    victims = [B200Overcooked("simple", 64, 0, horizon=400, seed=1), make_policies(lp, 1)[0]]
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ro._issue(False)
        for v in victims:
            v.close()
    g.replay()
    torch.cuda.synchronize()
    orc = COracle(lp, 256)
    o, r, d = orc.rollout(ro.buf.actions.cpu().numpy().astype(np.uint8))
    assert np.array_equal(ro.buf.obs[1:].cpu().numpy(), o) and np.array_equal(ro.buf.rewards.cpu().numpy(), r)


@pytest.mark.parametrize("use_graph", [False, True])
def test_crossplay_matrix_does_not_depend_on_the_sharding(use_graph):
    """with the sampling keyed by the GLOBAL (seat, world) (ocb_policy_set_sampling_rows) the
    per-pair returns of a sharded evaluation are bit-identical to the unsharded one — here 9 pairs in one env against
    three "ranks" of 3 pairs each on the same GPU."""
    layout, horizon, wpp, n_pol = "random1", 40, 128, 3
    lp = layouts.load_layout(layout, horizon)
    pol, _, _ = make_policies(lp, n_pol, gain=3.0)
    pairs = sharding.all_pairs(n_pol)
    total = len(pairs) * wpp

    def run(my_pairs, offset, with_total=True):
        ev = CrossPlayEvaluator(layout, pol, my_pairs, worlds_per_pair=wpp, horizon=horizon, seed=3, chunk_steps=20,
                                use_graph=use_graph, world_offset=offset, total_worlds=total if with_total else None)
        rs, ep = ev.run()
        torch.cuda.synchronize()
        out = (rs.clone(), ep.clone())
        ev.close()
        return out

    whole_rs, whole_ep = run(pairs, 0)
    parts = [run(sharding.pair_shard(pairs, r, 3), r * 3 * wpp) for r in range(3)]
    assert torch.equal(torch.cat([p[0] for p in parts]), whole_rs)
    assert torch.equal(torch.cat([p[1] for p in parts]), whole_ep)
    assert int(whole_rs.sum()) > 0
    # without the global rows the shards re-use the launch-local rows: same statistics, different streams
    local = [run(sharding.pair_shard(pairs, r, 3), r * 3 * wpp, with_total=False) for r in range(3)]
    assert not torch.equal(torch.cat([p[0] for p in local]), whole_rs)
    # and the unsharded run is the default stream when the env is the whole job (offset 0, N == total)
    plain_rs, _ = run(pairs, 0, with_total=False)
    assert torch.equal(plain_rs, whole_rs)
