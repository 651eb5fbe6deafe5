"""The drop-in claim on the GPU: the UNMODIFIED reference trainers (installed under baseline/_ref by
tools/install_reference.py, imported through oracle/ref_shim.py) drive ``B200Overcooked`` exactly as
train/trainer.py:13-44, train/testing.py:39-59 and train/XD/serial.py:25-75 wire them, and leave behind the same
bytes as on the reference's own ``SyncVectorEnv([SimplifiedOvercooked] * N)`` (device cuda, same torch seed):
every SharedReplayBuffer field after ``chooseinsert`` / ``diaginsert`` / ``partinsert``, episode scores, returns.

Also the other direction of the boundary: a device-collected ``PolicyRollout`` buffer, exposed with the reference's
13 ``SharedReplayBuffer`` fields, is consumed by the reference's ``compute_returns`` / ``feed_forward_generator`` /
``R_MAPPO.train`` unchanged.
"""
import numpy as np
import pytest
import torch

import trainer_harness as th
from diverse_conventions_b200 import layouts
from diverse_conventions_b200.env_utils import generate_env
from diverse_conventions_b200.overcooked_env import B200Overcooked
from diverse_conventions_b200.policy import FusedPolicy, PolicyNet
from diverse_conventions_b200.rollout import PolicyRollout

pytestmark = [pytest.mark.gpu, pytest.mark.reference]


@pytest.fixture(scope="module")
def ns():
    return th.load()


@pytest.mark.parametrize("layout,N,T,horizon", [("simple", 8, 14, 5), ("unident_s", 4, 8, 30)])
def test_main_player_with_centralized_partner(ns, tmp_path, layout, N, T, horizon):
    args = th.make_args(ns, layout, N, T, extra=["--cuda"])
    ref = th.run_main_player(ns, th.reference_env(ns, layout, N, horizon, "cuda"), "cuda", args, tmp_path / "ref")
    env = B200Overcooked(layout, N, 0, horizon=horizon)
    new = th.run_main_player(ns, env, "cuda", args, tmp_path / "new")
    # first episode: same initial weights, same sampling stream -> everything bit-equal
    a, b = ref["episodes"][0], new["episodes"][0]
    th.assert_buffers_equal(a["buffer"], b["buffer"], "episode 0")
    assert a["scores"] == b["scores"] and torch.equal(a["returns"], b["returns"])
    for k in a["train_infos"]:  # the PPO update saw identical minibatches (GPU reductions may differ in the last bits)
        assert a["train_infos"][k] == pytest.approx(b["train_infos"][k], rel=1e-3, abs=1e-6), k
    if horizon < T:
        assert (a["buffer"]["masks"][1:] == 0).any() and len(a["scores"]) > 0
    # the second episode ran on the updated weights through the same loop (run()'s body) and stayed well-formed
    b1 = new["episodes"][1]["buffer"]
    assert torch.isfinite(b1["value_preds"]).all() and b1["obs"].abs().sum() > 0
    if all(torch.equal(x, y) for x, y in zip(ref["actor"], new["actor"])):
        th.assert_buffers_equal(ref["episodes"][1]["buffer"], b1, "episode 1")


def test_generate_env_factory_and_run_sim(ns, tmp_path):
    args = th.make_args(ns, "simple", 6, 8, extra=["--cuda"])
    ref = th.run_sim_text(ns, th.reference_env(ns, "simple", 6, 200, "cuda"), "cuda", args, tmp_path / "ref")
    env = generate_env("overcooked", 6, "simple")  # the reference factory's signature (train/env_utils.py:10-28)
    new = th.run_sim_text(ns, env, "cuda", args, tmp_path / "new")
    assert ref == new and "STDEV" in ref


def test_xd_player_slices_and_mixed_play(ns, tmp_path):
    L, threads, horizon = 6, 3, 4
    args = th.make_args(ns, "simple", threads, L, extra=["--cuda", "--mp_weight", "0.5"])
    ref = th.run_xd_player(ns, lambda n: th.reference_env(ns, "simple", n, horizon, "cuda"), "cuda", args, tmp_path / "r",
                           threads)
    new = th.run_xd_player(ns, lambda n: B200Overcooked("simple", n, 0, horizon=horizon), "cuda", args, tmp_path / "n",
                           threads)
    for k in ("sp", "xp0", "xp1", "mp"):
        th.assert_buffers_equal(ref[k], new[k], k)
    assert ref["scores"] == new["scores"] and ref["mp_scores"] == new["mp_scores"] and ref["best_i"] == new["best_i"]
    assert any(len(s) for s in ref["scores"])


def test_device_rollout_feeds_the_reference_ppo_update(ns, tmp_path):
    """RolloutBuffer.shared_buffer_views / fill_shared_replay_buffer -> the reference's SharedReplayBuffer.compute_returns
    (shared_buffer.py:248-304), feed_forward_generator (:306-366) and R_MAPPO.train (r_mappo.py:166-224)"""
    layout, N, T, horizon = "simple", 128, 16, 6
    args = th.make_args(ns, layout, N, T, extra=["--cuda"])
    th.set_seed(3)
    env = B200Overcooked(layout, N, 0, horizon=horizon, seed=2)
    from pathlib import Path
    player = ns.MainPlayer({"all_args": args, "envs": env, "device": "cuda", "num_agents": 2, "run_dir": Path(tmp_path)})
    lp = layouts.load_layout(layout, horizon)
    pol = FusedPolicy(lp, 64, 1)
    actor = PolicyNet("actor", lp.width, lp.height, lp.channels, 64).load_state_dict(player.policy.actor.state_dict())
    critic = PolicyNet("critic", lp.width, lp.height, lp.channels, 64).load_state_dict(player.policy.critic.state_dict())
    pol.set_weights(0, actor, critic)
    ro = PolicyRollout(env, pol, T, seed=11)
    buf = ro.collect()
    torch.cuda.synchronize()

    views = buf.shared_buffer_views()
    sb = player.buffer
    for k in th.BUFFER_FIELDS:
        assert tuple(views[k].shape) == tuple(getattr(sb, k).shape), k
    buf.fill_shared_replay_buffer(sb)
    assert torch.equal(sb.obs, views["obs"].float()) and torch.equal(sb.masks[1:, :, 0, 0], (1 - buf.dones).float())
    assert (sb.masks[0] == 1).all() and (sb.masks[1:] == 0).any()

    # the reference's own value head on the stored observations agrees with the kernel's values (bf16x3 operands)
    with torch.no_grad():
        v_ref = player.policy.get_values(sb.share_obs[:-1].flatten(0, 2), sb.rnn_states_critic[:-1].flatten(0, 2),
                                         sb.masks[:-1].flatten(0, 2)).reshape(T, N, 2, 1)
    assert float((v_ref - sb.value_preds[:-1]).abs().max() / v_ref.abs().max()) < 1e-3

    # reference returns (its T-step Python loop, here as CUDA eager ops) == the gae kernel of this package (bit-identity
    # with the CPU reference is pinned by tests/test_gpu_returns.py; eager CUDA kernels may contract a*b+c differently)
    next_values = sb.value_preds[-1].clone()
    sb.compute_returns(next_values, player.trainer.value_normalizer)
    ours, _ = buf.compute_returns(args.gamma, args.gae_lambda, True, player.trainer.value_normalizer, normalize=False)
    ref_ret = sb.returns[..., 0]
    assert float((ref_ret - ours.transpose(1, 2)).abs().max()) <= 1e-5 * float(ref_ret.abs().max())

    adv = sb.returns[:-1] - player.trainer.value_normalizer.denormalize(sb.value_preds[:-1])
    batches = list(sb.feed_forward_generator(adv, num_mini_batch=2))
    assert len(batches) == 2 and batches[0][1].shape == (T * N * 2 // 2, lp.width, lp.height, lp.channels)
    player.trainer.prep_training()
    infos = player.trainer.train(sb)
    assert all(np.isfinite(float(v)) for v in infos.values()) and float(infos["ratio"]) == pytest.approx(1.0, abs=5e-3)
