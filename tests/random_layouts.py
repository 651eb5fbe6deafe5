"""Random Overcooked layouts for the differential tests (shared by tests/test_reference_random_layouts.py, the golden
generator tests/golden/make_random_layout_golden.py and the golden replays)."""
import numpy as np

NON_WALKABLE = "XXXXXXPODST"  # weights of the solid cell kinds


def random_layout(rng) -> dict:
    """a grid the CUDA path accepts (solid border, players on interior AIR) with at least one of every station"""
    while True:
        W, H = int(rng.integers(4, 10)), int(rng.integers(4, 7))
        g = [[NON_WALKABLE[int(rng.integers(len(NON_WALKABLE)))] for _ in range(W)] for _ in range(H)]
        air = []
        for y in range(1, H - 1):
            for x in range(1, W - 1):
                if rng.random() < 0.7:
                    g[y][x] = " "
                    air.append((x, y))
        n_players = int(rng.integers(1, 5))
        flat = "".join("".join(r) for r in g)
        if len(air) < n_players + 1 or not all(c in flat for c in "PODS"):
            continue
        for i, k in enumerate(rng.permutation(len(air))[:n_players]):
            x, y = air[int(k)]
            g[y][x] = str(i + 1)
        d = {"grid": "\n".join("".join(r) for r in g), "start_order_list": None}
        kind = int(rng.integers(3))
        if kind == 0:
            d["cook_time"], d["delivery_reward"] = int(rng.integers(1, 25)), int(rng.integers(1, 60))
        elif kind == 1:
            d["onion_time"], d["tomato_time"] = int(rng.integers(1, 9)), int(rng.integers(1, 9))
            d["onion_value"], d["tomato_value"] = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        d["rew_shaping_params"] = None if rng.random() < 0.5 else {
            "PLACEMENT_IN_POT_REW": int(rng.integers(0, 7)), "DISH_PICKUP_REWARD": int(rng.integers(0, 7)),
            "SOUP_PICKUP_REWARD": int(rng.integers(0, 9)), "DISH_DISP_DISTANCE_REW": 0, "POT_DISTANCE_REW": 0,
            "SOUP_DISTANCE_REW": 0}
        return d


def pack_ref_state(env) -> np.ndarray:
    """reference OvercookedState -> packed int32 row of include/ocb.h"""
    st = env.state
    P, S = env.mdp.num_players, env.mdp.size
    row = np.zeros(1 + 6 * P + 4 * S, dtype=np.int32)
    row[0] = st.timestep

    def put(at, obj):
        if obj != 0:
            row[at:at + 4] = (obj.name, obj.num_onions, obj.num_tomatoes, obj._cooking_tick)

    for i, pl in enumerate(st.players):
        row[1 + 6 * i] = pl.position
        row[1 + 6 * i + 1] = pl.orientation
        put(1 + 6 * i + 2, pl.held_object)
    for c in range(S):
        put(1 + 6 * P + 4 * c, st.objects[c])
    return row


def run_reference(ns, path, lp, horizon, steps, rng, team_cls):
    """the unmodified reference env on layout file `path` under scripted cooks with noise phases ->
    dict(actions [K,P], obs [K,P,W,H,C] int8, rewards [K], dones [K], states [K,L], reset_obs [P,W,H,C])"""
    import torch
    P = lp.num_players
    venv = ns.SyncVectorEnv([lambda: ns.SimplifiedOvercooked(path, horizon=horizon)], device="cpu")
    obs = venv.n_reset()
    env = venv.envs[0]
    team = team_cls(lp, rng, noise=0.2)
    out = {"actions": np.zeros((steps, P), np.uint8),
           "obs": np.zeros((steps, P, lp.width, lp.height, lp.channels), np.int8),
           "rewards": np.zeros((steps,), np.int32), "dones": np.zeros((steps,), np.int32),
           "states": np.zeros((steps, 1 + 6 * P + 4 * lp.size), np.int32),
           "reset_obs": np.stack([o.obs[0].numpy() for o in obs]).astype(np.int8)}
    for t in range(steps):
        team.noise = 1.0 if (t // 40) % 3 == 2 else 0.2
        a = np.asarray(team.joint(pack_ref_state(env)), dtype=np.int64)
        obs, r, dn, _ = venv.n_step(torch.from_numpy(a).reshape(P, 1, 1))
        o = np.stack([x.obs[0].numpy() for x in obs])
        assert np.array_equal(o, o.astype(np.int8))
        rr = r.numpy()
        assert np.all(rr == rr[0])  # team reward replicated per player
        out["actions"][t], out["obs"][t], out["rewards"][t], out["dones"][t] = a, o, int(rr[0, 0]), int(dn[0])
        out["states"][t] = pack_ref_state(env)
    return out
