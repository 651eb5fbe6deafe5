"""ctypes wrapper of oracle/ocb_oracle.c (TEST INFRASTRUCTURE ONLY, see that file)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libocb_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ocb_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ocb.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libocb_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        vp, i32, u64, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32
        L.ocbo_state_ints.argtypes = [vp]
        L.ocbo_state_ints.restype = i32
        L.ocbo_reset.argtypes = [vp, vp, i32]
        L.ocbo_observe.argtypes = [vp, vp, vp, i32]
        L.ocbo_step.argtypes = [vp, vp, vp, vp, vp, vp, i32]
        L.ocbo_rollout.argtypes = [vp, vp, i32, vp, vp, vp, vp, i32]
        L.ocbo_random_action.argtypes = [u64, u32, u64, i32, i32, i32]
        L.ocbo_random_action.restype = i32
        L.ocbo_random_actions.argtypes = [u64, u32, i32, u64, i32, i32, i32, vp]
        L.bbo_reset_world.argtypes = [u64, u32, vp]
        L.bbo_observe.argtypes = [vp, vp, i32]
        L.bbo_step.argtypes = [u64, vp, vp, vp, vp, vp, i32]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class COracle:
    """Same surface as oracle.overcooked_oracle.OvercookedOracle, C speed."""

    def __init__(self, params, num_worlds: int):
        self.params = params
        self.cfg = params.to_config()
        self._cfgp = ctypes.cast(ctypes.pointer(self.cfg), ctypes.c_void_p)
        self.N = int(num_worlds)
        self.P = params.num_players
        self.W, self.H, self.C = params.width, params.height, params.channels
        self.L = lib().ocbo_state_ints(self._cfgp)
        self.state = np.zeros((self.N, self.L), dtype=np.int32)
        self.reset()

    def reset(self):
        lib().ocbo_reset(self._cfgp, _p(self.state), self.N)

    def get_state(self):
        return self.state.copy()

    def set_state(self, st):
        self.state[:] = np.asarray(st, dtype=np.int32).reshape(self.N, self.L)

    def observe(self):
        obs = np.empty((self.P, self.N, self.W, self.H, self.C), dtype=np.int8)
        lib().ocbo_observe(self._cfgp, _p(self.state), _p(obs), self.N)
        return obs

    def step(self, actions, with_obs=True):
        a = np.ascontiguousarray(np.asarray(actions).reshape(self.P, self.N), dtype=np.int32)
        obs = np.empty((self.P, self.N, self.W, self.H, self.C), dtype=np.int8) if with_obs else None
        rew = np.empty((self.P, self.N), dtype=np.int32)
        done = np.empty((self.N,), dtype=np.int32)
        lib().ocbo_step(self._cfgp, _p(self.state), _p(a), _p(obs), _p(rew), _p(done), self.N)
        return obs, rew, done

    def rollout(self, actions_u8, with_obs=True):
        a = np.ascontiguousarray(actions_u8, dtype=np.uint8)
        K = a.shape[0]
        assert a.shape == (K, self.P, self.N)
        obs = np.empty((K, self.P, self.N, self.W, self.H, self.C), dtype=np.int8) if with_obs else None
        rew = np.empty((K, self.P, self.N), dtype=np.int32)
        done = np.empty((K, self.N), dtype=np.int32)
        lib().ocbo_rollout(self._cfgp, _p(self.state), K, _p(a), _p(obs), _p(rew), _p(done), self.N)
        return obs, rew, done


def random_actions(seed, world0, num_worlds, step0, num_steps, num_players, num_actions=6):
    out = np.empty((num_steps, num_players, num_worlds), dtype=np.uint8)
    lib().ocbo_random_actions(seed, world0, num_worlds, step0, num_steps, num_players, num_actions, _p(out))
    return out


class CBalanceOracle:
    def __init__(self, num_worlds: int, seed: int):
        self.N, self.seed = int(num_worlds), int(seed)
        self.state = np.zeros((self.N, 8), dtype=np.int32)
        self.reset()

    def reset(self):
        for n in range(self.N):
            row = self.state[n].copy()
            lib().bbo_reset_world(self.seed, n, _p(row))
            self.state[n] = row

    def observe(self):
        obs = np.empty((2, self.N, 7), dtype=np.int32)
        lib().bbo_observe(_p(self.state), _p(obs), self.N)
        return obs

    def step(self, actions):
        a = np.ascontiguousarray(np.asarray(actions).reshape(2, self.N), dtype=np.int32)
        obs = np.empty((2, self.N, 7), dtype=np.int32)
        rew = np.empty((2, self.N), dtype=np.float32)
        done = np.empty((self.N,), dtype=np.int32)
        lib().bbo_step(self.seed, _p(self.state), _p(a), _p(obs), _p(rew), _p(done), self.N)
        return obs, rew, done
