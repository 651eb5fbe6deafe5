"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol
include/ocb.h declares, fails loudly without a GPU, and the host mirror of the
reference's VectorMultiAgentEnv interface behaves like the reference's."""
import ctypes

import pytest
import torch

from diverse_conventions_b200 import _native, layouts
from diverse_conventions_b200.vector_api import (Discrete, MultiBinary, PlayerException, RandomVectorAgent,
                                                 VectorAgent, VectorMultiAgentEnv, VectorObservation)


def test_library_exports_every_declared_symbol():
    L = _native.lib()
    declared = _native.declared_symbols()
    assert len(declared) >= 30 and "ocb_step" in declared and "bb_step" in declared
    assert set(declared) == set(_native.PROTOTYPES)
    for name in declared:
        assert hasattr(L, name), name
    assert L.ocb_abi_version() == 1


def test_config_struct_matches_header_size():
    # uint32 + 7 int32 + 2*16 int32 + 2*4 int32 + 256 bytes
    assert ctypes.sizeof(layouts.ocb_config) == 4 * (8 + 32 + 8) + 256


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    L = _native.lib()
    cfg = layouts.load_layout("simple", 400).to_config()
    h = ctypes.c_void_p()
    rc = L.ocb_create(ctypes.byref(cfg), 0, 16, 0, ctypes.byref(h))
    assert rc < 0 and not h.value
    assert b"no CUDA device" in L.ocb_last_error() or b"cudaGetDeviceCount" in L.ocb_last_error()
    with pytest.raises(_native.NativeError):
        _native.check(rc)
    from diverse_conventions_b200.overcooked_env import B200Overcooked
    with pytest.raises(RuntimeError):
        B200Overcooked("simple", 4)


def test_create_rejects_bad_arguments():
    L = _native.lib()
    h = ctypes.c_void_p()
    cfg = layouts.load_layout("simple", 400).to_config()
    cfg.struct_size = 12
    assert L.ocb_create(ctypes.byref(cfg), 0, 16, 0, ctypes.byref(h)) < 0
    assert L.ocb_create(None, 0, 16, 0, ctypes.byref(h)) < 0
    assert L.ocb_num_worlds(None) < 0 and L.ocb_step(None, None, None, None, None, None) < 0


class _FakeEnv(VectorMultiAgentEnv):
    """2-player counter env used to exercise the ego-centric wrapper logic."""

    def __init__(self, n):
        super().__init__(n, torch.device("cpu"))
        self.observation_space = MultiBinary([2, 2, 3])
        self.action_space = Discrete(6)
        self.t = 0
        self.seen = None

    def _make_obs(self):
        return [VectorObservation(torch.ones(self.num_envs, dtype=torch.bool),
                                  torch.full((self.num_envs, 2, 2, 3), self.t + i, dtype=torch.int8)) for i in range(2)]

    def n_step(self, actions):
        self.seen = actions.clone()
        self.t += 1
        rew = torch.stack([actions[0, :, 0], actions[0, :, 0]]).to(torch.int32)
        return self._make_obs(), rew, torch.full((self.num_envs,), int(self.t % 3 == 0), dtype=torch.int32), [{}] * self.num_envs

    def n_reset(self):
        self.t = 0
        return self._make_obs()


class _Recorder(VectorAgent):
    def __init__(self, n):
        self.n, self.updates, self.obs = n, [], []

    def get_action(self, obs, record=True):
        self.obs.append(obs)
        return torch.full((self.n, 1), 4, dtype=torch.int64)

    def update(self, rewards, dones):
        self.updates.append((rewards.clone(), dones.clone()))


def test_vector_env_wrapper_semantics():
    env = _FakeEnv(5)
    partner = _Recorder(5)
    env.add_partner_agent(partner)
    ob = env.reset()
    assert ob.state is ob.obs and ob.obs.shape == (5, 2, 2, 3)
    ego = torch.arange(5).reshape(5, 1)
    ob, rew, done, info = env.step(ego)
    assert env.seen.shape == (2, 5, 1) and torch.equal(env.seen[0], ego) and torch.all(env.seen[1] == 4)
    assert torch.equal(rew, torch.arange(5, dtype=torch.int32)) and len(info) == 5
    assert len(partner.updates) == 1 and torch.equal(partner.updates[0][0], rew)
    assert int(partner.obs[0].obs[0, 0, 0, 0]) == 1  # the partner saw player 1's observation
    env.ego_ind = 1  # XD trainers flip the ego seat (train/XD/xd_player.py:668)
    env.step(ego)
    assert torch.equal(env.seen[1], ego) and torch.all(env.seen[0] == 4)


def test_vector_env_player_exceptions():
    env = _FakeEnv(2)
    with pytest.raises(PlayerException):
        env.add_partner_agent(RandomVectorAgent(lambda: None), player_num=0)
    with pytest.raises(PlayerException):
        env.set_resample_policy("bogus")
    with pytest.raises(PlayerException):
        VectorMultiAgentEnv.__init__(env, 2, "cpu", n_players=3, resample_policy="robin")
    with pytest.raises(PlayerException):
        VectorMultiAgentEnv.__init__(env, 2, "cpu", partners=[[], []])
    env2 = _FakeEnv(2)
    a, b = RandomVectorAgent(lambda: 0), RandomVectorAgent(lambda: 1)
    env2.add_partner_agent(a)
    env2.add_partner_agent(b)
    env2.reset()
    assert env2.partnerids == [1]
    env2.reset()
    assert env2.partnerids == [0]
    env2.set_partnerid(1)
    assert env2.partnerids == [1]
