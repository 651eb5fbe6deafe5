"""``B200BalanceBeam`` — drop-in for ``BalanceMadronaTorch`` (envs/balance_beam_env.py:22-41)
behind ``MadronaEnv.n_step`` / ``n_reset`` (pantheonrl_extension/vectorenv.py:306-343).

Two agents on a 5-cell beam, moves {-2,-1,+1,+2}, 3-step episodes
(semantics: PantheonLine, envs/balance_beam_env.py:95-152).  Observations int32 [N, 7],
state is the same tensor, action mask [N, 4] all-true, rewards float32 [2, N],
dones int32 [N].
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _native
from .vector_api import Discrete, MultiDiscrete, VectorMultiAgentEnv, VectorObservation

NUM_SPACES, BUFFER, TIME = 5, 2, 3
VALID_MOVES = [-2, -1, 1, 2]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class B200BalanceBeam(VectorMultiAgentEnv):
    def __init__(self, num_envs, gpu_id=0, debug_compile=True, use_cpu=False, use_env_cpu=False, seed=0):
        if use_cpu:
            raise RuntimeError("B200BalanceBeam has no CPU execution mode")
        if not torch.cuda.is_available():
            raise RuntimeError("B200BalanceBeam needs a CUDA device; there is no CPU fallback")
        self._lib = _native.lib()
        self.sim_device = torch.device("cuda", gpu_id)
        handle = ctypes.c_void_p()
        _native.check(self._lib.bb_create(gpu_id, num_envs, seed, ctypes.byref(handle)))
        self._h = handle
        super().__init__(num_envs, device=torch.device("cpu") if use_env_cpu else self.sim_device, n_players=2)
        N, dev = num_envs, self.sim_device
        self.static_observations = torch.empty((2, N, 2 * TIME + 1), dtype=torch.int32, device=dev)
        self.static_rewards = torch.zeros((2, N), dtype=torch.float32, device=dev)
        self.static_dones = torch.zeros((N,), dtype=torch.int32, device=dev)
        self.static_active_agents = torch.ones((2, N), dtype=torch.bool, device=dev)
        self.static_action_masks = torch.ones((N, len(VALID_MOVES)), dtype=torch.bool, device=dev)
        self.obs_size = self.state_size = 2 * TIME + 1
        self.discrete_action_size = len(VALID_MOVES)
        self.infos = [{}] * N
        self.observation_space = MultiDiscrete([NUM_SPACES + 2 * BUFFER] * 2 * TIME + [TIME])
        self.action_space = Discrete(len(VALID_MOVES))
        self.share_observation_space = self.observation_space
        self.observe()

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.sim_device).cuda_stream)

    def to_torch(self, a):
        return a if a.device == self.device else a.to(self.device)

    def close(self, **kwargs):
        if getattr(self, "_h", None):
            self._lib.bb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_obs(self):
        mask = self.to_torch(self.static_action_masks)
        out = []
        for i in range(2):
            o = self.to_torch(self.static_observations[i])
            out.append(VectorObservation(self.to_torch(self.static_active_agents[i]), o, o, mask))
        return out

    def observe(self):
        with torch.cuda.device(self.sim_device):
            _native.check(self._lib.bb_observe(self._h, _ptr(self.static_observations), self._stream()))
        return self.get_obs()

    def n_step(self, actions):
        a = actions.to(self.sim_device)
        if a.dim() == 3:
            a = a.squeeze(-1)
        a = a.to(torch.int32).contiguous()
        with torch.cuda.device(self.sim_device):
            _native.check(self._lib.bb_step(self._h, _ptr(a), _ptr(self.static_observations), _ptr(self.static_rewards),
                                            _ptr(self.static_dones), self._stream()))
        return self.get_obs(), self.to_torch(self.static_rewards), self.to_torch(self.static_dones), self.infos

    def n_reset(self):
        # MadronaEnv.n_reset (vectorenv.py:331-343) only re-reads the current observation
        return self.observe()

    def hard_reset(self):
        with torch.cuda.device(self.sim_device):
            _native.check(self._lib.bb_reset(self._h, _ptr(self.static_observations), self._stream()))
        return self.get_obs()

    def rollout_random(self, K, obs=True, actions=True, out=None):
        N, dev = self.num_envs, self.sim_device
        if out is None:
            out = self.alloc_rollout(K, obs, actions)
        with torch.cuda.device(self.sim_device):
            _native.check(self._lib.bb_rollout_random(self._h, K, _ptr(out["obs"]), _ptr(out["rewards"]), _ptr(out["dones"]),
                                                      _ptr(out["actions"]), self._stream()))
        return out

    def alloc_rollout(self, K, obs=True, actions=True):
        N, dev = self.num_envs, self.sim_device
        return {
            "obs": torch.empty((K, 2, N, 7), dtype=torch.int32, device=dev) if obs else None,
            "rewards": torch.empty((K, 2, N), dtype=torch.float32, device=dev),
            "dones": torch.empty((K, N), dtype=torch.int32, device=dev),
            "actions": torch.empty((K, 2, N), dtype=torch.uint8, device=dev) if actions else None,
        }

    def get_state(self):
        st = np.empty((self.num_envs, 8), dtype=np.int32)
        _native.check(self._lib.bb_get_state(self._h, st.ctypes.data_as(ctypes.c_void_p), st.size))
        return st

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=np.int32)
        _native.check(self._lib.bb_set_state(self._h, st.ctypes.data_as(ctypes.c_void_p), st.size))
