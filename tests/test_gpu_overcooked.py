"""GPU parity tests: the CUDA path, called through the C ABI / the drop-in adapter,
against the CPU oracle (oracle/ocb_oracle.c) and the committed golden vectors produced by
the reference's own Python env.  Bit-exact: states, rewards, dones, observations."""
import ctypes
import hashlib
import os

import numpy as np
import pytest
import torch

from diverse_conventions_b200 import _native, layouts
from oracle.c_oracle import COracle, random_actions

pytestmark = pytest.mark.gpu

CLASSIC = ["simple", "unident_s", "random1", "random0", "random3"]


def make_env(name, N, horizon=400, **kw):
    from diverse_conventions_b200.overcooked_env import B200Overcooked
    return B200Overcooked(name, N, 0, horizon=horizon, **kw)



def set_tuning_or_skip(env, G, tma):
    """lanes_per_world = 16 (role-split kernel) exists for two players and at most two pots"""
    try:
        env.set_tuning(G, bool(tma))
    except RuntimeError as exc:
        if G == 16 and "role-split" in str(exc):
            pytest.skip("role-split kernel does not serve this layout")
        raise

def stack_obs(vobs):
    return torch.stack([o.obs for o in vobs]).cpu().numpy()


# --------------------------------------------------------------------------- golden replays
@pytest.mark.parametrize("name", layouts.builtin_layout_names())
def test_replays_reference_trajectory_step_api(golden_dir, name):
    """n_step through the adapter, one launch per step, vs the reference's own trajectory."""
    g = np.load(os.path.join(golden_dir, "overcooked_%s.npz" % name))
    env = make_env(name, 3, int(g["horizon"]))
    P = env.num_players
    assert np.array_equal(stack_obs(env.n_reset())[:, 1], g["reset_obs"])
    T = 500 if name in CLASSIC else 450
    sha = hashlib.sha256()
    for t in range(T):
        a = torch.from_numpy(g["actions"][t].astype(np.int64)).reshape(P, 1, 1).repeat(1, 3, 1)
        vobs, rew, done, infos = env.n_step(a)
        o = stack_obs(vobs)
        assert vobs[0].state is vobs[0].obs and vobs[0].obs.dtype == torch.int8
        assert np.all(rew.cpu().numpy() == g["rewards"][t]), t
        assert np.all(done.cpu().numpy() == g["dones"][t]), t
        if t < g["obs_head"].shape[0]:
            assert np.array_equal(o[:, 2], g["obs_head"][t]), t
        assert np.array_equal(o[:, 0], o[:, 1]) and np.array_equal(o[:, 0], o[:, 2])
        sha.update(np.ascontiguousarray(o[:, 0]).tobytes())
    st = env.get_state()
    assert np.array_equal(st[0], g["states"][T - 1]) and np.array_equal(st[2], g["states"][T - 1])
    if T == g["actions"].shape[0]:
        assert sha.digest() == g["obs_sha256"].tobytes()


@pytest.mark.parametrize("name", CLASSIC + ["multiplayer_schelling", "simple_tomato", "simple_single"])
@pytest.mark.parametrize("G,tma", [(4, 0), (4, 1), (2, 0), (1, 1), (16, 1), (16, 0)])
def test_replays_reference_trajectory_fused(golden_dir, name, G, tma):
    """the same 1200 steps as ONE fused launch (ocb_rollout_actions); SHA over all observations.  G = 16 is the role-split
    kernel (transition warp + encoder warps; two players, at most two pots)"""
    g = np.load(os.path.join(golden_dir, "overcooked_%s.npz" % name))
    env = make_env(name, 5, int(g["horizon"]))
    set_tuning_or_skip(env, G, tma)
    T, P = g["actions"].shape
    a = torch.from_numpy(g["actions"]).reshape(T, P, 1).repeat(1, 1, 5).cuda()
    out = env.rollout_actions(a)
    torch.cuda.synchronize()
    obs = out["obs"].cpu().numpy()
    assert np.all(out["rewards"].cpu().numpy() == g["rewards"][:, None, None])
    assert np.all(out["dones"].cpu().numpy() == g["dones"][:, None])
    assert np.array_equal(obs[:64, :, 3], g["obs_head"])
    for n in (0, 4):
        assert hashlib.sha256(np.ascontiguousarray(obs[:, :, n]).tobytes()).digest() == g["obs_sha256"].tobytes()
    assert np.array_equal(env.get_state()[4], g["states"][-1])
    rs, ep = env.episode_stats()
    assert int(ep[0]) == 3 and int(rs[0]) == int(g["rewards"].sum())


# --------------------------------------------------------------------------- random play vs oracle
@pytest.mark.parametrize("name,N,horizon", [("simple", 1003, 37), ("unident_s", 257, 50), ("random0", 64, 400),
                                            ("corridor", 33, 60), ("multiplayer_schelling", 100, 45),
                                            ("simple_single", 9, 20), ("mdp_test", 130, 33)])
@pytest.mark.parametrize("G,tma", [(4, 0), (4, 1), (2, 1), (1, 0), (8, 1), (0, 1), (16, 1), (16, 0)])
def test_random_rollout_matches_oracle(name, N, horizon, G, tma):
    lp = layouts.load_layout(name, horizon)
    env = make_env(name, N, horizon, seed=1234)
    set_tuning_or_skip(env, G, tma)
    orc = COracle(lp, N)
    step0 = 0
    for K in (1, 7, 150):  # several launches, RNG stream continues across them
        out = env.rollout_random(K)
        torch.cuda.synchronize()
        acts = out["actions"].cpu().numpy()
        assert np.array_equal(acts, random_actions(1234, 0, N, step0, K, lp.num_players))
        o, r, d = orc.rollout(acts)
        assert np.array_equal(out["rewards"].cpu().numpy(), r)
        assert np.array_equal(out["dones"].cpu().numpy(), d)
        assert np.array_equal(out["obs"].cpu().numpy(), o)
        step0 += K
        assert env.step_count == step0
    assert np.array_equal(env.get_state(), orc.state)


@pytest.mark.parametrize("name,N,horizon,G,tile", [("simple", 1003, 37, 1, 28), ("simple", 1003, 37, 1, 7), ("simple", 1003, 37, 4, 7),
                                                   ("simple", 1003, 37, 16, 28), ("simple", 1003, 37, 16, 12), ("simple", 1003, 37, 16, 3),
                                                   ("unident_s", 515, 50, 2, 5), ("unident_s", 515, 50, 16, 20),
                                                   ("random0", 333, 40, 16, 9), ("multiplayer_schelling", 100, 45, 2, 13)])
def test_narrow_tiles_match_oracle(monkeypatch, name, N, horizon, G, tile):
    """K-step launches with tiles narrower than the warp (what `balanced_tile` picks to load every SM evenly; forced here
    through OCB_TILE_WORLDS, including widths that break the 16-byte alignment of the bulk stores)"""
    monkeypatch.setenv("OCB_TILE_WORLDS", str(tile))
    lp = layouts.load_layout(name, horizon)
    env = make_env(name, N, horizon, seed=99)
    set_tuning_or_skip(env, G, True)
    orc = COracle(lp, N)
    step0 = 0
    for K in (5, 90):
        out = env.rollout_random(K)
        torch.cuda.synchronize()
        acts = out["actions"].cpu().numpy()
        assert np.array_equal(acts, random_actions(99, 0, N, step0, K, lp.num_players))
        o, r, d = orc.rollout(acts)
        assert np.array_equal(out["rewards"].cpu().numpy(), r) and np.array_equal(out["dones"].cpu().numpy(), d)
        assert np.array_equal(out["obs"].cpu().numpy(), o)
        step0 += K
    assert np.array_equal(env.get_state(), orc.state)
    rs, ep = env.episode_stats()
    assert int(ep.sum()) == N * (step0 // horizon)


@pytest.mark.parametrize("name,N,lanes", [("simple", 4100, 0), ("random1", 4133, 0), ("unident_s", 4096, 0), ("simple_single", 4099, 0),
                                          ("random3", 300, 1), ("random0", 77, 2), ("multiplayer_schelling", 130, 4)])
def test_single_steps_match_oracle_across_episode_ends(name, N, lanes):
    """ONE env step per launch (n_step): from 4,096 worlds on the library serves a world with two lanes; the planes of a warp
    tile are rebuilt by one bulk copy of the template tile per view and only counters / pots are moved between HBM and shared
    memory.  Ragged world counts, an episode end inside the run (auto-reset + post-reset observation), every lane count."""
    horizon = 4
    lp = layouts.load_layout(name, horizon)
    env = make_env(name, N, horizon)
    if lanes:
        env.set_tuning(lanes, True)
    P = lp.num_players
    orc = COracle(lp, N)
    rng = np.random.default_rng(3)
    o0 = stack_obs(env.n_reset())
    assert np.array_equal(o0, orc.observe())
    for t in range(11):
        a = rng.integers(0, 6, size=(P, N))
        vobs, rew, done, _ = env.n_step(torch.from_numpy(a).reshape(P, N, 1))
        o, r, d = orc.step(a)
        assert np.array_equal(rew.cpu().numpy().astype(np.int64), r.astype(np.int64)), t
        assert np.array_equal(done.cpu().numpy().astype(np.int64), d.astype(np.int64)), t
        assert np.array_equal(stack_obs(vobs), o), t
        assert bool(d[0]) == ((t + 1) % horizon == 0)
    assert np.array_equal(env.get_state(), orc.state)


def test_scripted_teams_all_branches_many_worlds():
    """scripted cooks (+noise) on every classic layout: every reward branch, N worlds diverging"""
    from test_kernel_logic_emulated import scripted_actions
    for name in CLASSIC:
        lp = layouts.load_layout(name, 120)
        N, K = 96, 400
        acts = scripted_actions(lp, N, K, np.random.default_rng(11), noise=0.2)
        orc = COracle(lp, N)
        o, r, d = orc.rollout(acts)
        env = make_env(name, N, 120)
        out = env.rollout_actions(torch.from_numpy(acts).cuda())
        assert np.array_equal(out["rewards"].cpu().numpy(), r)
        assert np.array_equal(out["dones"].cpu().numpy(), d)
        assert np.array_equal(out["obs"].cpu().numpy(), o)
        vals = set(np.unique(r).tolist())
        assert {0, 3, 5, 20} <= vals, (name, vals)


# --------------------------------------------------------------------------- state injection
def test_state_injection_one_step():
    """reference protocol (validate_step, envs/balance_beam_env.py:157-218): inject states, step once, compare"""
    from test_kernel_logic_emulated import scripted_actions
    lp = layouts.load_layout("random3", 400)
    N = 512
    rng = np.random.default_rng(3)
    orc = COracle(lp, N)
    orc.rollout(scripted_actions(lp, N, 90, rng, noise=0.3))  # diverse reachable states
    env = make_env("random3", N)
    env.set_state(orc.state)
    assert np.array_equal(env.get_state(), orc.state)
    assert np.array_equal(stack_obs(env.observe()), orc.observe())
    for _ in range(5):
        a = rng.integers(0, 6, size=(2, N))
        o, r, d = orc.step(a)
        vobs, rew, done, _ = env.n_step(torch.from_numpy(a).reshape(2, N, 1))
        assert np.array_equal(stack_obs(vobs), o) and np.array_equal(rew.cpu().numpy(), r)
        assert np.array_equal(done.cpu().numpy(), d) and np.array_equal(env.get_state(), orc.state)


def test_set_state_rejects_unrepresentable_states():
    env = make_env("simple", 4)
    st = env.get_state()
    bad = st.copy()
    bad[1, 1] = 0  # player 0 standing on a counter
    with pytest.raises(_native.NativeError) as e:
        env.set_state(bad)
    assert e.value.code == -5
    bad = st.copy()
    cell0 = 1 + 6 * 2
    bad[2, cell0 + 4 * 6: cell0 + 4 * 6 + 4] = (3, 0, 0, -1)  # dish floating on an AIR cell
    with pytest.raises(_native.NativeError):
        env.set_state(bad)
    bad = st.copy()
    bad[3, 1 + 6] = bad[3, 1]  # both players on one cell: unreachable in the MDP, the plane update assumes it never happens
    with pytest.raises(_native.NativeError):
        env.set_state(bad)
    assert np.array_equal(env.get_state(), st)  # rejected states leave the env untouched
    with pytest.raises(_native.NativeError):
        env.set_state(st[:2])


def test_graph_replays_of_random_rollouts_continue_the_action_stream():
    """a captured ocb_rollout_random reads its first step from the device counter: replays draw fresh actions, the
    host mirror (ocb_step_count) follows, and an uncaptured launch afterwards continues where the replays stopped"""
    N, K, seed = 96, 8, 77
    env = make_env("simple", N, seed=seed)
    out = env.alloc_rollout(K)
    env.rollout_random(K, out)  # uncaptured warm-up: steps 0..7
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        env.rollout_random(K, out)
    acts = []
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        acts.append(out["actions"].cpu().numpy().copy())
    assert env.step_count == 4 * K
    assert np.array_equal(np.concatenate(acts), random_actions(seed, 0, N, K, 3 * K, 2))
    tail = env.rollout_random(K)
    torch.cuda.synchronize()
    assert np.array_equal(tail["actions"].cpu().numpy(), random_actions(seed, 0, N, 4 * K, K, 2))
    assert env.step_count == 5 * K


# --------------------------------------------------------------------------- API conventions
def test_action_dtypes_devices_and_invalid_actions():
    lp = layouts.load_layout("simple", 400)
    N = 50
    rng = np.random.default_rng(0)
    a = rng.integers(0, 6, size=(2, N, 1))
    ref = None
    for dt, dev in [(torch.int64, "cpu"), (torch.int32, "cuda"), (torch.float32, "cuda"), (torch.uint8, "cuda"),
                    (torch.int16, "cpu"), (torch.float64, "cuda")]:
        env = make_env("simple", N)
        vobs, rew, done, _ = env.n_step(torch.from_numpy(a).to(dt).to(dev))
        cur = (stack_obs(vobs), rew.cpu().numpy())
        if ref is None:
            orc = COracle(lp, N)
            o, r, d = orc.step(a[..., 0])
            ref = (o, r)
        assert np.array_equal(cur[0], ref[0]) and np.array_equal(cur[1], ref[1]), dt
    env = make_env("simple", N)
    before = env.get_state()
    env.n_step(torch.full((2, N, 1), 17, dtype=torch.int64))  # out of range -> STAY
    after = env.get_state()
    before[:, 0] += 1
    assert np.array_equal(before, after)
    with pytest.raises(ValueError):
        env.n_step(torch.zeros((3, N, 1)))


def test_adapter_contract_and_ego_wrapper():
    from diverse_conventions_b200.vector_api import RandomVectorAgent
    env = make_env("random1", 32, horizon=10)
    assert env.observation_space.shape == (5, 5, 20) and env.action_space.n == 6
    assert env.action_space.__class__.__name__ == "Discrete" and env.share_observation_space is env.observation_space
    assert env.num_envs == 32 and env.n_players == 2 and env.device.type == "cuda"
    env.add_partner_agent(RandomVectorAgent(lambda: torch.randint(0, 6, (32, 1), device="cuda").float()))
    ob = env.reset()
    assert ob.obs.shape == (32, 5, 5, 20) and ob.action_mask.shape == (32, 6) and ob.action_mask.all()
    assert ob.active.dtype == torch.bool and ob.active.all()
    total = torch.zeros(32, device="cuda")
    for t in range(10):
        ob, rew, done, info = env.step(torch.randint(0, 6, (32, 1), device="cuda").float())
        total += rew
        assert rew.shape == (32,) and done.shape == (32,) and len(info) == 32
        assert bool(done.to(torch.bool).all()) == (t == 9)
    # done step returns the post-reset observation
    fresh = make_env("random1", 32, horizon=10)
    assert torch.equal(ob.obs, fresh.n_reset()[0].obs)
    env_cpu = make_env("simple", 8, use_env_cpu=True)
    vobs, rew, done, _ = env_cpu.n_step(torch.zeros((2, 8, 1)))
    assert vobs[0].obs.device.type == "cpu" and rew.device.type == "cpu"


def test_step_host_matches_device_path():
    N = 200
    env_a, env_b = make_env("simple", N), make_env("simple", N)
    rng = np.random.default_rng(1)
    h_act = torch.empty((2, N), dtype=torch.int32).pin_memory()
    h_obs = torch.empty((2, N, 5, 4, 20), dtype=torch.int8).pin_memory()
    h_rew = torch.empty((2, N), dtype=torch.int32).pin_memory()
    h_done = torch.empty((N,), dtype=torch.int32).pin_memory()
    for t in range(30):
        h_act.copy_(torch.from_numpy(rng.integers(0, 6, size=(2, N)).astype(np.int32)))
        env_a.step_host(h_act, h_obs, h_rew, h_done)
        vobs, rew, done, _ = env_b.n_step(h_act.reshape(2, N, 1))
        assert torch.equal(torch.stack([o.obs for o in vobs]).cpu(), h_obs)
        assert torch.equal(rew.cpu(), h_rew) and torch.equal(done.cpu(), h_done)


def test_step_host_async_pipeline_matches_oracle():
    """two steps in flight (ocb_step_host_async / ocb_step_host_wait): every retired step's host buffers equal the oracle's
    step, in order, including across an episode end; mixing with the synchronous call afterwards keeps the order"""
    N, horizon, T = 333, 9, 25
    lp = layouts.load_layout("random1", horizon)
    env = make_env("random1", N, horizon)
    orc = COracle(lp, N)
    rng = np.random.default_rng(3)
    acts = rng.integers(0, 6, size=(T, 2, N)).astype(np.int32)
    bufs = [dict(a=torch.empty((2, N), dtype=torch.int32).pin_memory(),
                 o=torch.empty((2, N, lp.width, lp.height, lp.channels), dtype=torch.int8).pin_memory(),
                 r=torch.empty((2, N), dtype=torch.int32).pin_memory(), d=torch.empty((N,), dtype=torch.int32).pin_memory())
            for _ in range(2)]
    retired = 0

    def retire():
        nonlocal retired
        left = env.step_host_wait()
        b = bufs[retired & 1]
        o, r, d = orc.step(acts[retired])
        assert np.array_equal(b["o"].numpy(), o) and np.array_equal(b["r"].numpy(), r) and np.array_equal(b["d"].numpy(), d)
        retired += 1
        return left

    for t in range(T):
        if t >= 2:
            retire()  # the slot about to be reused
        b = bufs[t & 1]
        b["a"].copy_(torch.from_numpy(acts[t]))
        env.step_host_async(b["a"], b["o"], b["r"], b["d"])
    assert retire() == 1 and retire() == 0 and env.step_host_wait() == 0
    assert retired == T and env.step_count == T
    a = rng.integers(0, 6, size=(2, N)).astype(np.int32)
    bufs[0]["a"].copy_(torch.from_numpy(a))
    env.step_host(bufs[0]["a"], bufs[0]["o"], bufs[0]["r"], bufs[0]["d"])
    o, r, d = orc.step(a)
    assert np.array_equal(bufs[0]["o"].numpy(), o) and np.array_equal(env.get_state(), orc.state)


def test_error_codes_through_the_abi():
    L = _native.lib()
    h = ctypes.c_void_p()
    lp = layouts.load_layout("simple", 400)
    cfg = lp.to_config()
    cfg.terrain[6] = 9
    assert L.ocb_create(ctypes.byref(cfg), 0, 4, 0, ctypes.byref(h)) == -2 and b"terrain" in L.ocb_last_error()
    cfg = lp.to_config()
    assert L.ocb_create(ctypes.byref(cfg), 99, 4, 0, ctypes.byref(h)) == -1
    assert L.ocb_create(ctypes.byref(cfg), 0, 0, 0, ctypes.byref(h)) == -1
    assert L.ocb_create(ctypes.byref(cfg), 0, 4, 0, ctypes.byref(h)) == 0
    assert L.ocb_set_tuning(h, 3, 0) == -1 and L.ocb_step(h, None, None, None, None, None) == -1
    assert L.ocb_num_players(h) == 2 and L.ocb_obs_bytes_per_agent(h) == 400 and L.ocb_obs_channels(h) == 20
    assert L.ocb_state_ints_per_world(h) == 1 + 12 + 80
    assert L.ocb_destroy(h) == 0


# --------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("N,lanes", [(16384, 0), (8192, 0), (16384, 16)])
def test_full_size_properties_cramped_room(N, lanes):
    """BASELINE config 3 size (16,384 worlds, horizon 400; default launch shape = the one-warp kernel) and config 4's
    8,192 worlds (default = the role-split kernel), plus the split kernel forced at 16,384: size-independent properties +
    an oracle replay of a strided subset of worlds."""
    H, K = 400, 100
    lp = layouts.load_layout("simple", H)
    env = make_env("simple", N, H, seed=7)
    if lanes:
        env.set_tuning(lanes, True)
    assert env.get_tuning()["lanes_per_world"] == (16 if (lanes == 16 or N < 16384) else 1)
    sub = np.arange(0, N, 331)
    orc = COracle(lp, len(sub))
    total_rew = torch.zeros((), dtype=torch.int64, device="cuda")
    n_done = 0
    out = env.alloc_rollout(K)
    for it in range(10):  # 1000 steps = 2.5 episodes
        env.rollout_random(K, out)
        obs, rew, done, acts = out["obs"], out["rewards"], out["dones"], out["actions"]
        assert torch.equal(rew[:, 0], rew[:, 1])  # team reward replicated
        total_rew += rew[:, 0].sum()
        n_done += int(done.sum())
        steps = torch.arange(it * K + 1, (it + 1) * K + 1, device="cuda")
        assert torch.equal(done.bool(), ((steps % H) == 0)[:, None].expand(K, N))  # exactly every H steps
        # structural invariants of the encoding: one cell per player channel, static terrain channels
        assert int(obs[:, :, :, :, :, 0].sum()) == K * 2 * N and int(obs[:, :, :, :, :, 1].sum()) == K * 2 * N
        assert int(obs[:, :, :, :, :, 2:10].sum()) == K * 2 * N * 2
        assert torch.equal(obs[:, 0, :, :, :, 10:15], obs[:, 1, :, :, :, 10:15])
        assert torch.equal(obs[0, 0, 0, :, :, 10:15].expand(K, N, 5, 4, 5), obs[:, 0, :, :, :, 10:15])
        assert torch.equal(obs[:, 0, :, :, :, 0], obs[:, 1, :, :, :, 1])  # views mirror each other
        a = acts.cpu().numpy()
        o, r, d = orc.rollout(np.ascontiguousarray(a[:, :, sub]))
        assert np.array_equal(obs[:, :, sub].cpu().numpy(), o)
        assert np.array_equal(rew[:, :, sub].cpu().numpy(), r)
    assert n_done == 2 * N
    rs, ep = env.episode_stats()
    assert int(ep.sum()) == 2 * N
    st = env.get_state()
    assert np.all(st[:, 0] == 200) and np.array_equal(st[sub], orc.state)
    # device-side return accounting == sum of rewards of completed episodes + the running one
    assert int(rs.sum()) <= int(total_rew)


@pytest.mark.parametrize("name", CLASSIC)
def test_config4_size_env_rollouts_match_oracle(name):
    """the five classic layouts at BASELINE config 4's 8,192 worlds (default launch shape: the role-split kernel for the
    small planes, the one-warp kernel for asymmetric_advantages / counter_circuit): 450 random steps across an episode end,
    a strided subset of worlds replayed through the oracle, every world's done flags and the episode accounting checked"""
    N, H, K = 8192, 400, 90
    lp = layouts.load_layout(name, H)
    env = make_env(name, N, H, seed=11)
    sub = np.arange(3, N, 257)
    orc = COracle(lp, len(sub))
    out = env.alloc_rollout(K)
    for it in range(5):
        env.rollout_random(K, out)
        steps = torch.arange(it * K + 1, (it + 1) * K + 1, device="cuda")
        assert torch.equal(out["dones"].bool(), ((steps % H) == 0)[:, None].expand(K, N))
        assert torch.equal(out["rewards"][:, 0], out["rewards"][:, 1])
        a = out["actions"].cpu().numpy()
        o, r, d = orc.rollout(np.ascontiguousarray(a[:, :, sub]))
        assert np.array_equal(out["obs"][:, :, sub].cpu().numpy(), o)
        assert np.array_equal(out["rewards"][:, :, sub].cpu().numpy(), r)
    assert np.array_equal(env.get_state()[sub], orc.state)
    rs, ep = env.episode_stats()
    assert int(ep.sum()) == N


def test_observe_matches_rollout_last_obs_and_reset_idempotent():
    env = make_env("unident_s", 300, horizon=400, seed=5)
    out = env.rollout_random(40)
    assert torch.equal(torch.stack([o.obs for o in env.observe()]), out["obs"][-1])
    a = torch.stack([o.obs for o in env.n_reset()]).clone()
    b = torch.stack([o.obs for o in env.n_reset()]).clone()
    assert torch.equal(a, b)
    st = env.get_state()
    assert np.all(st[:, 0] == 0) and np.all(st[:, 13:] == 0)
