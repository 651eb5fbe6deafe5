#!/usr/bin/env python
"""Generate the committed golden vectors from the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference ships no golden vectors for the Overcooked path, so parity
is pinned on what its own code produces here:

  layouts.json            get_base_layout_params() of every shipped layout
                          (envs/overcooked2_env.py:171-291), max_num_players None/1/2
  overcooked_<name>.npz   one 1200-step trajectory per layout through
                          SyncVectorEnv([SimplifiedOvercooked(name, horizon=400)])
                          (pantheonrl_extension/vectorenv.py:348-425,
                          envs/overcooked2_env.py:294-343): actions, per-step rewards /
                          dones, packed state after every step, full observations of the
                          first 64 steps and a SHA-256 over all observations
  kat.json                KAT-1 / KAT-2 of SURVEY.md 8c (reset observation of
                          cramped_room, 43-step scripted soup)
  balance_beam.npz        every reachable PantheonLine transition
                          (envs/balance_beam_env.py:95-152) by state injection
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402
from diverse_conventions_b200 import layouts  # noqa: E402
from scripted_agent import ScriptedTeam  # noqa: E402

HORIZON = 400
STEPS = 1200
OBS_HEAD = 64


def pack_ref_state(env) -> np.ndarray:
    """Reference OvercookedState -> packed int32 row of include/ocb.h."""
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


def gen_layouts(ns):
    out = {}
    for name in layouts.builtin_layout_names():
        for mp in (None, 1, 2):
            out["%s|%s" % (name, mp)] = ns.get_base_layout_params(name, HORIZON, max_num_players=mp)
    with open(os.path.join(HERE, "layouts.json"), "w") as f:
        json.dump(out, f, sort_keys=True)
    print("layouts.json", len(out))


def gen_trajectory(ns, name, seed):
    params = layouts.load_layout(name, HORIZON)
    P = params.num_players
    rng = np.random.default_rng(seed)
    venv = ns.SyncVectorEnv([lambda: ns.SimplifiedOvercooked(name, horizon=HORIZON)], device="cpu")
    obs = venv.n_reset()
    env = venv.envs[0]
    team = ScriptedTeam(params, rng, noise=0.15)
    acts = np.zeros((STEPS, P), np.uint8)
    rews = np.zeros((STEPS,), np.int32)
    dones = np.zeros((STEPS,), np.uint8)
    states = np.zeros((STEPS, 1 + 6 * P + 4 * params.size), np.int32)
    head = np.zeros((OBS_HEAD, P, params.width, params.height, params.channels), np.int8)
    reset_obs = np.stack([o.obs[0].numpy() for o in obs]).astype(np.int8)
    sha = hashlib.sha256()
    for t in range(STEPS):
        # phases of pure noise exercise collisions / blocked moves / odd interacts
        team.noise = 1.0 if (t // 100) % 4 == 3 else 0.15
        a = np.asarray(team.joint(pack_ref_state(env)), dtype=np.int64)
        obs, r, d, _ = venv.n_step(torch.from_numpy(a).reshape(P, 1, 1))
        o8 = np.stack([o.obs[0].numpy() for o in obs])
        assert np.array_equal(o8, o8.astype(np.int8)), "observation is not int8-valued"
        o8 = np.ascontiguousarray(o8.astype(np.int8))
        rr = r.numpy()
        assert np.all(rr == rr[0]) and float(rr[0, 0]).is_integer()
        acts[t], rews[t], dones[t] = a, int(rr[0, 0]), int(d[0])
        states[t] = pack_ref_state(env)
        if t < OBS_HEAD:
            head[t] = o8
        sha.update(o8.tobytes())
    np.savez_compressed(os.path.join(HERE, "overcooked_%s.npz" % name), actions=acts, rewards=rews, dones=dones,
                        states=states, obs_head=head, reset_obs=reset_obs,
                        obs_sha256=np.frombuffer(sha.digest(), dtype=np.uint8), horizon=np.int32(HORIZON))
    hist = {int(k): int(v) for k, v in zip(*np.unique(rews, return_counts=True))}
    print(name, "P=%d" % P, "reward histogram", hist)


def gen_kat(ns):
    env = ns.SimplifiedOvercooked("simple", horizon=HORIZON)
    _, obs = env.n_reset()
    kat1 = []
    for p in range(2):
        o = obs[p][0]
        nz = np.argwhere(o != 0)
        kat1.append([[int(x), int(y), int(c), int(o[x, y, c])] for x, y, c in nz])
    a0 = [0, 3, 5, 2, 0, 5, 3, 3, 5, 2, 0, 5, 3, 3, 5, 2, 0, 5, 3, 1, 1, 5, 0, 0, 2, 0, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 5,
          1, 2, 1, 5]
    sha = hashlib.sha256()
    trace = []
    for a in a0:
        _, obs, rew, done, _ = env.n_step((torch.tensor(a), torch.tensor(4)))
        trace.append(int(rew[0]))
        for p in range(2):
            sha.update(np.ascontiguousarray(obs[p][0].astype(np.int8)).tobytes())
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump({"kat1_reset_nonzero": kat1, "kat2_actions_p0": a0, "kat2_action_p1": 4, "kat2_rewards": trace,
                   "kat2_obs_sha256": sha.hexdigest()}, f)
    print("kat2 total", sum(trace), sha.hexdigest())


def gen_balance():
    PantheonLine = ref_shim.load_balance()
    env = PantheonLine()
    rows = []

    def inject(x, y):
        env.n_reset()
        env.state = np.array([x, y])
        env.ego_state = np.zeros(3)
        env.alt_state = np.zeros(3)
        env.current_time = 2
        env.update_states()

    def snap():
        o = env.get_full_obs()
        return np.concatenate([o[0][0], o[1][0]]).astype(np.int32)

    for x in range(5):
        for y in range(5):
            for a1 in range(16):
                inject(x, y)
                pre = snap()
                _, _, rew, done, _ = env.n_step([[a1 // 4], [a1 % 4]])
                rows.append(np.concatenate([pre, [a1 // 4, a1 % 4], snap(), [np.float32(rew[0]).view(np.int32), int(done)]]))
                if done:
                    continue
                for a2 in range(16):
                    inject(x, y)
                    env.n_step([[a1 // 4], [a1 % 4]])
                    pre = snap()
                    _, _, rew, done, _ = env.n_step([[a2 // 4], [a2 % 4]])
                    rows.append(np.concatenate([pre, [a2 // 4, a2 % 4], snap(),
                                                [np.float32(rew[0]).view(np.int32), int(done)]]))
    rows = np.asarray(rows, dtype=np.int32)
    # columns: pre obs(14) | actions(2) | post obs (14, pre-reset) | reward fp32 bits | done
    np.savez_compressed(os.path.join(HERE, "balance_beam.npz"), transitions=rows)
    print("balance_beam transitions", rows.shape)


def main():
    ns = ref_shim.load()
    gen_layouts(ns)
    gen_kat(ns)
    for i, name in enumerate(layouts.builtin_layout_names()):
        gen_trajectory(ns, name, 1000 + i)
    gen_balance()


if __name__ == "__main__":
    main()
