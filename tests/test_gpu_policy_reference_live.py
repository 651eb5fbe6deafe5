"""The tensor-core policy kernels against the reference's OWN networks, live on the GPU box (markers gpu + reference; the
reference travels as baseline/_ref): R_Actor / R_Critic (train/MAPPO/r_actor_critic.py:12-71,142-197) are built by the
reference's code with its orthogonal init at hidden_size 64 (every train/*.sh) and 512 (the argparse default,
train/config.py:199), their state dicts are loaded into the kernels, and the kernels' logits / values are compared with the
reference modules' forward on the same observations.  Bar of the north star: 1e-3 relative; asserted at 2e-4."""
import numpy as np
import pytest
import torch

from diverse_conventions_b200 import layouts
from diverse_conventions_b200.policy import FusedPolicy, PolicyNet
from oracle.c_oracle import COracle

pytestmark = [pytest.mark.gpu, pytest.mark.reference]

REL_TOL = 2e-4


def played_obs(lp, N, steps=60, seed=2):
    orc = COracle(lp, N)
    rng = np.random.default_rng(seed)
    for _ in range(steps):
        o, _, _ = orc.step(rng.choice(6, size=(2, N), p=[.15, .15, .15, .15, .05, .35]))
    return torch.from_numpy(o.reshape(2 * N, lp.width, lp.height, lp.channels).copy())


@pytest.mark.parametrize("hidden", [64, 512])
@pytest.mark.parametrize("layout,N", [("simple", 300), ("random1", 131), ("unident_s", 97)])
def test_kernels_match_the_reference_modules(layout, N, hidden):
    from oracle import ref_shim
    R_Actor, R_Critic = ref_shim.load_policy()
    from config import get_config  # reference train/config.py
    import gym
    args = get_config().parse_args(["--hidden_size", str(hidden), "--gain", "1.0"])
    lp = layouts.load_layout(layout, 400)
    space = gym.spaces.MultiBinary([lp.width, lp.height, lp.channels])
    torch.manual_seed(7)
    actor, critic = R_Actor(args, space, gym.spaces.Discrete(6)), R_Critic(args, space)
    with torch.no_grad():  # non-zero biases: the reference initialises them to 0, training moves them
        for m in (actor, critic):
            for name, prm in m.named_parameters():
                if name.endswith("bias"):
                    prm.uniform_(-0.1, 0.1)
    obs = played_obs(lp, N)
    M = obs.shape[0]
    with torch.no_grad():
        ref_logits = actor.act.action_out.linear(actor.base(obs.float()))
        ref_values, _ = critic(obs, torch.zeros(M, 1, hidden), torch.ones(M, 1))
    pol = FusedPolicy(lp, hidden, 1)
    pol.set_weights(0, PolicyNet("actor", lp.width, lp.height, lp.channels, hidden).load_state_dict(actor.state_dict()),
                    PolicyNet("critic", lp.width, lp.height, lp.channels, hidden).load_state_dict(critic.state_dict()))
    out = pol.forward(obs.cuda(), deterministic=True, want_logits=True)
    torch.cuda.synchronize()
    err_l = float((out["logits"].cpu() - ref_logits).abs().max() / ref_logits.abs().max())
    err_v = float((out["values"].cpu() - ref_values[:, 0]).abs().max() / ref_values.abs().max())
    assert err_l < REL_TOL and err_v < REL_TOL, (err_l, err_v)
    assert torch.equal(out["actions"].cpu().long(), out["logits"].cpu().argmax(-1))
