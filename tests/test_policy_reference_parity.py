"""PolicyNet (the plain PyTorch fp32 forward the GPU tests compare against) versus the reference's own
R_Actor / R_Critic (train/MAPPO/r_actor_critic.py) at hidden sizes 64 and 512.  Runs wherever the reference is
reachable: /root/reference in the build container, the travelled install baseline/_ref on the GPU box."""
import os

import numpy as np
import pytest
import torch

from diverse_conventions_b200 import layouts
from diverse_conventions_b200.policy import PolicyNet

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("hidden", [64, 512])
@pytest.mark.parametrize("layout", ["simple", "random1"])
def test_policynet_equals_reference_networks(layout, hidden):
    from oracle import ref_shim
    R_Actor, R_Critic = ref_shim.load_policy()
    from config import get_config
    import gym
    args = get_config().parse_args([])
    args.hidden_size = hidden
    lp = layouts.load_layout(layout, 400)
    space = gym.spaces.MultiBinary([lp.width, lp.height, lp.channels])
    torch.manual_seed(3)
    actor, critic = R_Actor(args, space, gym.spaces.Discrete(6)), R_Critic(args, space)
    obs = torch.from_numpy(np.random.default_rng(0).integers(0, 2, size=(37, lp.width, lp.height, lp.channels)).astype(np.float32))
    mine_a = PolicyNet("actor", lp.width, lp.height, lp.channels, hidden).load_state_dict(actor.state_dict())
    mine_c = PolicyNet("critic", lp.width, lp.height, lp.channels, hidden).load_state_dict(critic.state_dict())
    with torch.no_grad():
        feats = actor.base(obs)
        ref_logits = actor.act.action_out.linear(feats)
        rnn = torch.zeros(37, 1, hidden)
        ref_values, _ = critic(obs, rnn, torch.ones(37, 1))
    assert torch.allclose(mine_a.forward(obs), ref_logits, rtol=1e-5, atol=1e-6)
    assert torch.allclose(mine_c.forward(obs), ref_values, rtol=1e-5, atol=1e-6)
