"""``OracleOvercooked`` — TEST-ONLY twin of ``B200Overcooked`` that runs on the CPU oracle (oracle/ocb_oracle.c).

It exists so that the drop-in contract (the reference's trainers driving this package's ``VectorMultiAgentEnv``
through the adapter's dtype / view conventions) is also checked by the ``-m "not gpu"`` suite: same static buffers
(int8 observations ``[P, N, W, H, C]``, int32 rewards ``[P, N]``, int32 dones ``[N]``, bool masks), same action
handling (``[P, N, 1]`` of any numeric dtype, truncated to int32 like ``load_action`` in csrc/oc_device.cuh), same
"returned tensors are views the next step overwrites" behaviour as overcooked_env.py.  Not a product path: the
product has no CPU simulator.
"""
import numpy as np
import torch

from diverse_conventions_b200 import layouts
from diverse_conventions_b200.vector_api import Discrete, MultiBinary, VectorMultiAgentEnv, VectorObservation
from oracle.c_oracle import COracle


class OracleOvercooked(VectorMultiAgentEnv):
    def __init__(self, layout_name, num_envs, horizon=200, ego_agent_idx=0, num_players=None):
        self.layout = layouts.load_layout(layout_name, horizon, num_players)
        self.oracle = COracle(self.layout, num_envs)
        self.width, self.height, self.channels = self.layout.width, self.layout.height, self.layout.channels
        self.num_players = self.layout.num_players
        super().__init__(num_envs, device=torch.device("cpu"), ego_ind=ego_agent_idx, n_players=self.num_players)
        P, N = self.num_players, num_envs
        self.static_observations = torch.zeros((P, N, self.width, self.height, self.channels), dtype=torch.int8)
        self.static_rewards = torch.zeros((P, N), dtype=torch.int32)
        self.static_dones = torch.zeros((N,), dtype=torch.int32)
        self.static_active_agents = torch.ones((P, N), dtype=torch.bool)
        self.static_action_mask = torch.ones((N, 6), dtype=torch.bool)
        self.infos = [{}] * N
        self.observation_space = MultiBinary(np.array([self.width, self.height, self.channels]))
        self.share_observation_space = self.observation_space
        self.action_space = Discrete(6)
        self.n_reset()

    def get_obs(self):
        return [VectorObservation(self.static_active_agents[i], self.static_observations[i],
                                  action_mask=self.static_action_mask) for i in range(self.n_players)]

    def n_step(self, actions):
        a = actions.detach().cpu()
        if a.dim() == 3:
            a = a.squeeze(-1)
        a = a.to(torch.int32).numpy()  # float -> int truncation, as the kernel's load_action
        obs, rew, done = self.oracle.step(a)
        self.static_observations.copy_(torch.from_numpy(obs))
        self.static_rewards.copy_(torch.from_numpy(rew))
        self.static_dones.copy_(torch.from_numpy(done))
        return self.get_obs(), self.static_rewards, self.static_dones, self.infos

    def n_reset(self):
        self.oracle.reset()
        self.static_observations.copy_(torch.from_numpy(self.oracle.observe()))
        return self.get_obs()
