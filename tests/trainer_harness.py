"""Drive the UNMODIFIED reference trainers on a given ``VectorMultiAgentEnv`` (test infrastructure).

The functions below wire the reference's classes exactly as its entry scripts do and return what they left
behind, so that a test can run them once on the reference's own ``SyncVectorEnv([SimplifiedOvercooked] * N)`` and
once on this package's env (``B200Overcooked`` on the GPU, ``tests/oracle_env.OracleOvercooked`` on the CPU) with
the same torch seed and compare everything bit for bit:

* ``run_main_player``  train/trainer.py:13-44  (MainPlayer + CentralizedAgent partner, ``ego.run()``'s loop:
  setup_data / warmup / collect_episode -> chooseinsert / compute / train), train/MAPPO/main_player.py:91-112,211-277
* ``run_sim_text``     train/testing.py:39-59 (``run_sim`` with two DecentralizedAgents)
* ``run_xd_player``    train/XD/serial.py:25-75 (XDPlayer + CentralizedMultiAgent slices + MixedAgent mixed play),
  train/XD/xd_player.py:85-230,232-356
"""
import contextlib
import io
import random
from pathlib import Path

import numpy as np
import torch

from oracle import ref_shim

BUFFER_FIELDS = ("share_obs", "obs", "rnn_states", "rnn_states_critic", "value_preds", "returns", "available_actions",
                 "actions", "action_log_probs", "rewards", "masks", "bad_masks", "active_masks")


def load():
    return ref_shim.load_trainers()


def reference_env(ns, layout, num_envs, horizon, device):
    """the reference's CPU env path: SyncVectorEnv over N SimplifiedOvercooked (pantheonrl_extension/vectorenv.py:348-425)"""

    class IntActionEnv(ns.SimplifiedOvercooked):
        # the trainers hand FLOAT action tensors to envs.step (turn_actions is fp32, main_player.py:268); the
        # reference's GPU adapter casts them on copy (envs/overcooked2_env.py:119), its Python env indexes a list
        # with act.item() and needs the same cast
        def n_step(self, actions):
            return super().n_step([a.to(torch.int64) for a in actions])

    return ns.SyncVectorEnv([lambda: IntActionEnv(layout, horizon=horizon) for _ in range(num_envs)], device=device)


def make_args(ns, layout, n_rollout_threads, episode_length, hidden_size=64, seed=1, extra=()):
    argv = ["--env_name", "overcooked", "--over_layout", layout, "--n_rollout_threads", str(n_rollout_threads),
            "--episode_length", str(episode_length), "--hidden_size", str(hidden_size), "--seed", str(seed),
            "--ppo_epoch", "2", "--num_mini_batch", "2", "--env_length", str(episode_length)]
    args = ns.get_config().parse_args(argv + list(extra))
    args.hanabi_name = layout
    return args


def set_seed(seed):
    torch.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)


def snapshot(buf):
    out = {}
    for k in BUFFER_FIELDS:
        v = getattr(buf, k, None)
        if v is not None:
            out[k] = v.detach().cpu().clone()
    return out


def run_main_player(ns, envs, device, args, run_dir, episodes=2):
    """trainer.py:13-44 + the body of MainPlayer.run (main_player.py:182-209)"""
    set_seed(args.seed)
    config = {"all_args": args, "envs": envs, "device": device, "num_agents": 2, "run_dir": Path(run_dir)}
    ego = ns.MainPlayer(config)
    partner = ns.CentralizedAgent(ego, 1)
    envs.add_partner_agent(partner)
    ego.setup_data()
    ego.warmup()
    out = {"episodes": []}
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(episodes):
            ego.collect_episode()
            ep = {"buffer": snapshot(ego.buffer), "scores": list(ego.scores)}
            ego.compute()
            ep["returns"] = ego.buffer.returns.detach().cpu().clone()
            ep["train_infos"] = {k: float(v.detach() if torch.is_tensor(v) else v) for k, v in ego.train().items()}
            out["episodes"].append(ep)
    out["actor"] = [p.detach().cpu().clone() for p in ego.policy.actor.parameters()]
    out["player"] = ego
    return out


def run_sim_text(ns, envs, device, args, run_dir):
    """testing.py:39-59: two DecentralizedAgents of freshly initialised MainPlayers, 200 steps; run_sim only prints"""
    set_seed(args.seed)
    players = []
    for k in range(2):
        config = {"all_args": args, "envs": envs, "device": device, "num_agents": 2, "run_dir": Path(run_dir) / str(k)}
        players.append(ns.MainPlayer(config))
    ego, alt = ns.DecentralizedAgent(players[0], 0), ns.DecentralizedAgent(players[1], 1)
    text = io.StringIO()
    with contextlib.redirect_stdout(text), torch.no_grad():
        ns.run_sim(envs, ego, alt)
    return text.getvalue()


def run_xd_player(ns, make_env, device, args, run_dir, threads):
    """XD/serial.py:25-75 for convention 1 of a population (one earlier convention in agent_set): self-play slice +
    two cross-play slices through CentralizedMultiAgent, then the mixed-play collection on envs_mp"""
    from XD.serial import generate_buffer
    agent_num = 1
    env = make_env(threads * (agent_num * 2 + 1))
    env_mp = make_env(args.env_length - 1)
    set_seed(args.seed)
    prior = ns.MCPolicy(args, env.observation_space, env.share_observation_space, env.action_space, 0, torch.device(device))
    set_seed(args.seed + int(args.seed_skip))
    pol = ns.MCPolicy(args, env.observation_space, env.share_observation_space, env.action_space, agent_num, torch.device(device))
    sp_buf = generate_buffer(args, env, device)
    xp_buf0 = [generate_buffer(args, env, device) for _ in range(agent_num)]
    xp_buf1 = [generate_buffer(args, env, device) for _ in range(agent_num)]
    mp_buf = generate_buffer(args, env_mp, device, args.env_length - 1)
    config = {"all_args": args, "envs": env, "envs_mp": env_mp, "device": device, "num_agents": 2, "run_dir": Path(run_dir)}
    with contextlib.redirect_stdout(io.StringIO()):
        runner = ns.XDPlayer(config, pol, sp_buf, xp_buf0, xp_buf1, mp_buf, [prior.actor], args.xp_weight, args.mp_weight,
                             args.mix_prob, args.env_length)
        runner.setup_data()
        runner.warmup()
        runner.collect_episode()
    return {"sp": snapshot(sp_buf), "xp0": snapshot(xp_buf0[0]), "xp1": snapshot(xp_buf1[0]), "mp": snapshot(mp_buf),
            "scores": [list(s) for s in runner.scores], "mp_scores": list(runner.mp_scores), "best_i": runner.best_i}


def assert_buffers_equal(a, b, what=""):
    assert a.keys() == b.keys(), (what, a.keys(), b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, (what, k, a[k].shape, b[k].shape)
        assert torch.equal(a[k], b[k]), "%s: SharedReplayBuffer.%s differs" % (what, k)
