"""Batched multi-agent env / agent interface — the drop-in boundary.

Mirrors the interface of the reference's ``pantheonrl_extension`` package so that its
trainers (train/MAPPO/main_player.py, train/partner_agents.py, train/XD/xd_player.py,
train/ADAP) can drive the B200 simulator unchanged:

* ``VectorObservation``      pantheonrl_extension/vectorobservation.py:5-32
* ``VectorAgent``            pantheonrl_extension/vectoragent.py:9-29
* ``RandomVectorAgent``      pantheonrl_extension/vectoragent.py:32-40
* ``VectorMultiAgentEnv``    pantheonrl_extension/vectorenv.py:26-255
* ``PlayerException``        pantheonrl_extension/vectorenv.py:13-14

Same class / method / attribute names, argument meaning and error behaviour; the code is
written against that contract, it is not a copy.
"""
from __future__ import annotations

import random
from abc import ABC, abstractmethod
from typing import List, Optional

import torch

try:  # the reference uses gym.spaces; consumers only read .shape, .n and the class name
    from gym.spaces import Discrete, MultiBinary, MultiDiscrete  # type: ignore
except Exception:  # gym is not installed on the B200 image
    class _Space:
        shape: tuple = ()

    class Discrete(_Space):  # noqa: D401  (class *name* matters: train/MAPPO/utils/act.py:18)
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()

    class MultiBinary(_Space):
        def __init__(self, n):
            self.n = n
            self.shape = tuple(int(x) for x in n)

    class MultiDiscrete(_Space):
        def __init__(self, nvec):
            self.nvec = list(nvec)
            self.shape = (len(self.nvec),)


class PlayerException(Exception):
    """Raised when the players of an environment are set up inconsistently."""


class VectorObservation:
    """Batched observation of one agent over N worlds.

    active [N] bool, obs [N, *obs_shape], state [N, *state_shape] (defaults to ``obs``,
    i.e. ``state is obs``), action_mask [N, num_actions] bool or None.
    """
    __slots__ = ("active", "obs", "state", "action_mask")

    def __init__(self, active: torch.Tensor, obs: torch.Tensor, state: Optional[torch.Tensor] = None,
                 action_mask: Optional[torch.Tensor] = None):
        self.active = active
        self.obs = obs
        self.state = obs if state is None else state
        self.action_mask = action_mask

    def __repr__(self):
        return "VectorObservation(obs=%s, state_is_obs=%s)" % (tuple(self.obs.shape), self.state is self.obs)


class VectorAgent(ABC):
    @abstractmethod
    def get_action(self, obs: VectorObservation, record: bool = True) -> torch.Tensor:
        """Action [N, 1] for the given batched observation."""

    @abstractmethod
    def update(self, rewards: torch.Tensor, dones: torch.Tensor) -> None:
        """Reward [N] / done [N] feedback for the most recent recorded action."""


class RandomVectorAgent(VectorAgent):
    def __init__(self, sampler):
        self.sampler = sampler

    def get_action(self, obs: VectorObservation, record: bool = True) -> torch.Tensor:
        return self.sampler()

    def update(self, rewards: torch.Tensor, dones: torch.Tensor) -> None:
        return None


class VectorMultiAgentEnv(ABC):
    """Ego-centric wrapper around a batched N-world, P-player simulator.

    ``step(ego_action)`` asks the registered partner agents for the other players'
    actions, stacks everything to [P, N, 1], calls ``n_step`` and hands rewards / dones
    back to the partners; it returns the ego player's view.
    """

    def __init__(self, num_envs: int, device, ego_ind: int = 0, n_players: int = 2, resample_policy: str = "default",
                 partners: Optional[List[List[VectorAgent]]] = None):
        self.num_envs = num_envs
        self.device = device
        self.ego_ind = ego_ind
        self.n_players = n_players
        if partners is not None:
            if len(partners) != n_players - 1:
                raise PlayerException("The number of partners needs to equal the number of non-ego players")
            for plist in partners:
                if not isinstance(plist, list) or not plist:
                    raise PlayerException("Sublist for each partner must be nonempty list")
        self.partners = partners or [[] for _ in range(n_players - 1)]
        self.partnerids = [0] * (n_players - 1)
        self._obs = tuple()
        self._actions = None
        self.set_resample_policy(resample_policy)

    # -- partner management -------------------------------------------------------
    def getDummyEnv(self, player_num: int):
        return self

    def _get_partner_num(self, player_num: int) -> int:
        if player_num == self.ego_ind:
            raise PlayerException("Ego agent is not set by the environment")
        return player_num - 1 if player_num > self.ego_ind else player_num

    def add_partner_agent(self, agent: VectorAgent, player_num: int = 1) -> None:
        self.partners[self._get_partner_num(player_num)].append(agent)

    def set_partnerid(self, agent_id: int, player_num: int = 1) -> None:
        partner_num = self._get_partner_num(player_num)
        assert 0 <= agent_id < len(self.partners[partner_num])
        self.partnerids[partner_num] = agent_id

    def resample_random(self) -> None:
        self.partnerids = [random.randrange(len(plist)) if plist else 0 for plist in self.partners]

    def resample_round_robin(self) -> None:
        self.partnerids = [(self.partnerids[0] + 1) % max(len(self.partners[0]), 1)]

    def set_resample_policy(self, resample_policy: str) -> None:
        if resample_policy == "default":
            resample_policy = "robin" if self.n_players == 2 else "random"
        if resample_policy == "robin" and self.n_players != 2:
            raise PlayerException("Cannot do round robin resampling for >2 players")
        if resample_policy == "robin":
            self.resample_partner = self.resample_round_robin
        elif resample_policy == "random":
            self.resample_partner = self.resample_random
        else:
            raise PlayerException(f"Invalid resampling policy: {resample_policy}")

    # -- stepping -----------------------------------------------------------------
    def _get_actions(self, obs, ego_act=None):
        actions = []
        for player, ob in zip(range(self.n_players), obs):
            if player == self.ego_ind:
                actions.append(ego_act)
            else:
                p = self._get_partner_num(player)
                actions.append(self.partners[p][self.partnerids[p]].get_action(ob))
        if self._actions is None or self._actions.shape[1:] != actions[0].shape or self._actions.dtype != actions[0].dtype:
            self._actions = torch.stack(actions)
        else:
            torch.stack(actions, out=self._actions)
        return self._actions

    def _update_players(self, rews, done):
        for i in range(self.n_players - 1):
            playernum = i + (0 if i < self.ego_ind else 1)
            self.partners[i][self.partnerids[i]].update(rews[playernum], done)

    def step(self, action: torch.Tensor):
        acts = self._get_actions(self._obs, action)
        self._obs, rews, done, info = self.n_step(acts)
        self._update_players(rews, done)
        return self._obs[self.ego_ind], rews[self.ego_ind], done, info

    def reset(self):
        self.resample_partner()
        self._obs = self.n_reset()
        return self._obs[self.ego_ind]

    @abstractmethod
    def n_step(self, actions: torch.Tensor):
        """actions [P, N, 1] -> (List[VectorObservation], rewards [P, N], dones [N], infos)."""

    @abstractmethod
    def n_reset(self):
        """-> List[VectorObservation], one per player."""

    def close(self, **kwargs):
        pass
