#!/usr/bin/env python
"""Golden vectors for the policy forward, generated from the reference's own R_Actor / R_Critic
(train/MAPPO/r_actor_critic.py) in the build container:

  policy_<layout>_h64.npz : random-init (torch seed 1, the reference's orthogonal init, gain 0.01)
      actor + critic state dicts, 384 observations taken from a scripted trajectory of the
      reference env, the actor's pre-softmax logits (R_Actor.get_logits(...).logits is the
      normalised version; we store act.action_out.linear(features)) and the critic's values.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402
from oracle.c_oracle import COracle  # noqa: E402
from diverse_conventions_b200 import layouts  # noqa: E402
from scripted_agent import ScriptedTeam  # noqa: E402


def main():
    R_Actor, R_Critic = ref_shim.load_policy()
    from config import get_config  # reference train/config.py
    import gym
    for layout in ("simple", "random1"):
        args = get_config().parse_args([])
        args.hidden_size = 64
        lp = layouts.load_layout(layout, 400)
        space = gym.spaces.MultiBinary([lp.width, lp.height, lp.channels])
        torch.manual_seed(1)
        actor = R_Actor(args, space, gym.spaces.Discrete(6))
        critic = R_Critic(args, space)
        # observations: scripted play so that pots / dishes / soups appear
        N = 64
        orc = COracle(lp, N)
        team = ScriptedTeam(lp, np.random.default_rng(5), noise=0.2)
        rows = []
        for t in range(120):
            acts = np.array([team.joint(orc.state[n]) for n in range(N)]).T
            o, _, _ = orc.step(acts)
            if t % 20 == 19:
                rows.append(o.reshape(-1, lp.width, lp.height, lp.channels))
        obs = np.concatenate(rows)[:384]
        x = torch.from_numpy(obs)
        with torch.no_grad():
            M = x.shape[0]
            rnn, masks = torch.zeros(M, 1, 64), torch.ones(M, 1)
            dist = actor.get_logits(x, rnn, masks, torch.ones(M, 6))
            feat = actor.base(x.float())
            raw = actor.act.action_out.linear(feat)
            assert torch.allclose(dist.logits, torch.log_softmax(raw, -1), atol=1e-6)
            values, _ = critic(x, rnn, masks)
        out = {"obs": obs, "logits": raw.numpy(), "values": values.numpy()}
        for k, v in actor.state_dict().items():
            out["actor." + k] = v.numpy()
        for k, v in critic.state_dict().items():
            out["critic." + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, "policy_%s_h64.npz" % layout), **out)
        print(layout, obs.shape, "logits range", float(raw.min()), float(raw.max()), "values", float(values.min()), float(values.max()))


if __name__ == "__main__":
    main()
