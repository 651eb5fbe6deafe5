"""``B200Overcooked`` — drop-in for the reference's ``OvercookedMadrona`` env adapter.

Reference: envs/overcooked2_env.py:27-131 (``OvercookedMadrona(VectorMultiAgentEnv)``),
which glues the Madrona simulator to torch with gather / index_put kernels and a host
sync per step.  Here ``n_step`` is ONE kernel launch on torch's current stream that
writes observations, rewards and dones straight into torch-owned tensors in their
final ``[P, N, W, H, C]`` / ``[P, N]`` / ``[N]`` layouts; nothing synchronises.

Conventions kept from the reference:
  * actions ``[P, N, 1]`` of any numeric dtype / device;
  * ``VectorObservation(active [N] bool, obs [N, W, H, C] int8, state is obs,
    action_mask [N, 6] bool all-true)`` per player;
  * rewards ``[P, N]`` int32 (team reward replicated), dones ``[N]`` int32;
  * on a done step the observation is the post-reset one;
  * the returned tensors are views of static buffers that the next step overwrites
    (consumers clone, train/MAPPO/main_player.py:245-247).
``n_reset`` really resets every world (as the Python oracle's SyncVectorEnv does;
the Madrona adapter's n_reset only re-reads the current observation).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import numpy as np
import torch

from . import _native
from .layouts import LayoutParams, load_layout, io_bytes_per_world_step
from .vector_api import Discrete, MultiBinary, VectorMultiAgentEnv, VectorObservation

NUM_ACTIONS = 6

_DTYPE_CODES = {torch.int32: _native.ACT_I32, torch.int64: _native.ACT_I64, torch.float32: _native.ACT_F32,
                torch.uint8: _native.ACT_U8}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class B200Overcooked(VectorMultiAgentEnv):
    def __init__(self, layout_name, num_envs, gpu_id=0, debug_compile=True, use_cpu=False, use_env_cpu=False,
                 ego_agent_idx=0, horizon=200, num_players=None, seed=0, world_offset=0,
                 layout_params: Optional[LayoutParams] = None):
        if use_cpu:
            raise RuntimeError("B200Overcooked has no CPU execution mode (use the reference's Python env)")
        if not torch.cuda.is_available():
            raise RuntimeError("B200Overcooked needs a CUDA device; there is no CPU fallback")
        self._lib = _native.lib()
        self.layout_name = layout_name
        self.layout = layout_params if layout_params is not None else load_layout(layout_name, horizon, num_players)
        self.base_layout_params = self.layout.as_dict()
        self.width, self.height = self.layout.width, self.layout.height
        self.num_players = self.layout.num_players
        self.size = self.width * self.height
        self.channels = self.layout.channels
        self.horizon = horizon
        self.sim_device = torch.device("cuda", gpu_id)

        self._cfg = self.layout.to_config()
        handle = ctypes.c_void_p()
        _native.check(self._lib.ocb_create(ctypes.byref(self._cfg), gpu_id, num_envs, seed, ctypes.byref(handle)))
        self._h = handle
        if world_offset:
            _native.check(self._lib.ocb_set_world_offset(self._h, world_offset))

        env_device = torch.device("cpu") if use_env_cpu else self.sim_device
        super().__init__(num_envs, device=env_device, ego_ind=ego_agent_idx, n_players=self.num_players)

        P, N = self.num_players, num_envs
        dev = self.sim_device
        self.static_observations = torch.empty((P, N, self.width, self.height, self.channels), dtype=torch.int8, device=dev)
        self.static_rewards = torch.zeros((P, N), dtype=torch.int32, device=dev)
        self.static_dones = torch.zeros((N,), dtype=torch.int32, device=dev)
        self.static_active_agents = torch.ones((P, N), dtype=torch.bool, device=dev)
        self.static_action_mask = torch.ones((N, NUM_ACTIONS), dtype=torch.bool, device=dev)
        self._actions_i32 = torch.empty((P, N), dtype=torch.int32, device=dev)
        self._p_obs, self._p_rew, self._p_done = (_ptr(self.static_observations), _ptr(self.static_rewards),
                                                  _ptr(self.static_dones))

        self.obs_size = self.size * self.channels
        self.state_size = self.obs_size
        self.infos = [{}] * N
        self.ego_ind = ego_agent_idx
        self.observation_space = MultiBinary(np.array([self.width, self.height, self.channels]))
        self.share_observation_space = self.observation_space
        self.action_space = Discrete(NUM_ACTIONS)
        self.io_bytes_per_world_step = io_bytes_per_world_step(self.layout)
        self.n_reset()

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.sim_device).cuda_stream)

    def to_torch(self, a):
        return a if a.device == self.device else a.to(self.device)

    def close(self, **kwargs):
        if getattr(self, "_h", None):
            self._lib.ocb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tuning(self, lanes_per_world: int = 0, use_tma: bool = False):
        _native.check(self._lib.ocb_set_tuning(self._h, lanes_per_world, int(use_tma)))

    def get_tuning(self) -> dict:
        g, t, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _native.check(self._lib.ocb_get_tuning(self._h, ctypes.byref(g), ctypes.byref(t), ctypes.byref(w)))
        return {"lanes_per_world": g.value, "use_tma": bool(t.value), "warps_per_cta": w.value}

    # ------------------------------------------------------------------ reference API
    def get_obs(self) -> List[VectorObservation]:
        if self.device == self.sim_device:  # views of the static buffers, sliced once (tensor indexing costs microseconds)
            views = self.__dict__.get("_static_views")
            if views is None:
                views = self._static_views = [(self.static_active_agents[i], self.static_observations[i])
                                              for i in range(self.n_players)]
            mask = self.static_action_mask
            return [VectorObservation(a, o, action_mask=mask) for a, o in views]
        mask = self.to_torch(self.static_action_mask)
        return [VectorObservation(self.to_torch(self.static_active_agents[i]), self.to_torch(self.static_observations[i]),
                                  action_mask=mask) for i in range(self.n_players)]

    def _prepare_actions(self, actions: torch.Tensor, lead: int):
        """-> (tensor on the sim device, dtype code); [lead..., P, N(,1)] -> contiguous [lead..., P, N]"""
        a = actions
        if a.device != self.sim_device:
            a = a.to(self.sim_device, non_blocking=True)
        if a.dim() == lead + 3:
            a = a.squeeze(-1)
        if a.shape[-2:] != (self.num_players, self.num_envs):
            raise ValueError("actions must have shape [..., %d, %d(, 1)], got %s" %
                             (self.num_players, self.num_envs, tuple(actions.shape)))
        code = _DTYPE_CODES.get(a.dtype)
        if code is None:
            a = a.to(torch.int32)
            code = _native.ACT_I32
        return a.contiguous(), code

    def n_step(self, actions: torch.Tensor):
        a, code = self._prepare_actions(actions, 0)
        # the library switches to the env's device itself (DeviceGuard); torch only has to name the right stream
        rc = self._lib.ocb_step_ex(self._h, _ptr(a), code, self._p_obs, self._p_rew, self._p_done, self._stream())
        if rc < 0:
            _native.check(rc)
        if self.device == self.sim_device:
            return self.get_obs(), self.static_rewards, self.static_dones, self.infos
        return self.get_obs(), self.to_torch(self.static_rewards), self.to_torch(self.static_dones), self.infos

    def n_reset(self):
        with torch.cuda.device(self.sim_device):
            _native.check(self._lib.ocb_reset(self._h, _ptr(self.static_observations), self._stream()))
        return self.get_obs()

    # ------------------------------------------------------------------ beyond the reference API
    def observe(self) -> List[VectorObservation]:
        with torch.cuda.device(self.sim_device):
            _native.check(self._lib.ocb_observe(self._h, _ptr(self.static_observations), self._stream()))
        return self.get_obs()

    def alloc_rollout(self, K: int, obs: bool = True, actions: bool = True):
        P, N, dev = self.num_players, self.num_envs, self.sim_device
        return {
            "obs": torch.empty((K, P, N, self.width, self.height, self.channels), dtype=torch.int8, device=dev) if obs else None,
            "rewards": torch.empty((K, P, N), dtype=torch.int32, device=dev),
            "dones": torch.empty((K, N), dtype=torch.int32, device=dev),
            "actions": torch.empty((K, P, N), dtype=torch.uint8, device=dev) if actions else None,
        }

    def rollout_random(self, K: int, out: Optional[dict] = None, obs: bool = True, actions: bool = True) -> dict:
        """K fused steps with on-device uniform random actions (ocb_rollout_random)."""
        out = out if out is not None else self.alloc_rollout(K, obs, actions)
        with torch.cuda.device(self.sim_device):
            _native.check(self._lib.ocb_rollout_random(self._h, K, _ptr(out.get("obs")), _ptr(out.get("rewards")),
                                                       _ptr(out.get("dones")), _ptr(out.get("actions")), self._stream()))
        return out

    def rollout_actions(self, actions: torch.Tensor, out: Optional[dict] = None, obs: bool = True) -> dict:
        """K fused steps with caller supplied actions [K, P, N] (ocb_rollout_actions)."""
        a, code = self._prepare_actions(actions, 1)
        K = a.shape[0]
        out = out if out is not None else self.alloc_rollout(K, obs, False)
        with torch.cuda.device(self.sim_device):
            _native.check(self._lib.ocb_rollout_actions(self._h, K, _ptr(a), code, _ptr(out.get("obs")),
                                                        _ptr(out.get("rewards")), _ptr(out.get("dones")), self._stream()))
        return out

    def step_host(self, h_actions: torch.Tensor, h_obs=None, h_rewards=None, h_dones=None):
        """ocb_step_host: host (pinned) buffers in and out, synchronous."""
        assert h_actions.dtype == torch.int32 and h_actions.device.type == "cpu" and h_actions.is_contiguous()
        _native.check(self._lib.ocb_step_host(self._h, _ptr(h_actions), _ptr(h_obs), _ptr(h_rewards), _ptr(h_dones)))

    def step_host_async(self, h_actions: torch.Tensor, h_obs=None, h_rewards=None, h_dones=None):
        """ocb_step_host_async: enqueue one step of the two-deep host-buffer pipeline and return (pinned buffers)"""
        assert h_actions.dtype == torch.int32 and h_actions.device.type == "cpu" and h_actions.is_contiguous()
        _native.check(self._lib.ocb_step_host_async(self._h, _ptr(h_actions), _ptr(h_obs), _ptr(h_rewards), _ptr(h_dones)))

    def step_host_wait(self) -> int:
        """ocb_step_host_wait: block until the oldest enqueued step delivered its buffers -> steps still in flight"""
        rc = self._lib.ocb_step_host_wait(self._h)
        if rc < 0:
            _native.check(rc)
        return rc

    def get_state(self) -> np.ndarray:
        L = self._lib.ocb_state_ints_per_world(self._h)
        st = np.empty((self.num_envs, L), dtype=np.int32)
        _native.check(self._lib.ocb_get_state(self._h, st.ctypes.data_as(ctypes.c_void_p), st.size))
        return st

    def set_state(self, st) -> None:
        st = np.ascontiguousarray(st, dtype=np.int32)
        _native.check(self._lib.ocb_set_state(self._h, st.ctypes.data_as(ctypes.c_void_p), st.size))

    def episode_stats(self):
        """(sum of returns of completed episodes [N] int64, completed episodes [N] int32), on the device"""
        rs = torch.empty((self.num_envs,), dtype=torch.int64, device=self.sim_device)
        ep = torch.empty((self.num_envs,), dtype=torch.int32, device=self.sim_device)
        with torch.cuda.device(self.sim_device):
            _native.check(self._lib.ocb_read_episode_stats(self._h, _ptr(rs), _ptr(ep), self._stream()))
        return rs, ep

    def clear_episode_stats(self):
        with torch.cuda.device(self.sim_device):
            _native.check(self._lib.ocb_clear_episode_stats(self._h, self._stream()))

    @property
    def step_count(self) -> int:
        return int(self._lib.ocb_step_count(self._h))
