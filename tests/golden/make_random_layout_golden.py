#!/usr/bin/env python
"""Golden trajectories of the UNMODIFIED reference env on RANDOM layouts (beyond the 21 files it ships), so that the
GPU box — which has no reference checkout — can replay them through the CUDA path.  For each seed: a random grid
(tests/random_layouts.py) is written as a .layout file, parsed by the reference's get_base_layout_params and run
through SyncVectorEnv([SimplifiedOvercooked]) under scripted cooks with noise phases.

Writes tests/golden/random_layouts.npz: per seed k the layout dict (repr), horizon, parsed parameters (json), actions,
rewards, dones, the packed state after every step, the reset observation and a SHA-256 over all observations."""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from diverse_conventions_b200 import layouts  # noqa: E402
from oracle import ref_shim  # noqa: E402
from random_layouts import random_layout, run_reference  # noqa: E402
from scripted_agent import ScriptedTeam  # noqa: E402

SEEDS = list(range(100, 110))
STEPS = 300


def main():
    ns = ref_shim.load()
    out = {"seeds": np.asarray(SEEDS)}
    tmp = tempfile.mkdtemp()
    for seed in SEEDS:
        rng = np.random.default_rng(seed)
        d = random_layout(rng)
        path = os.path.join(tmp, "rand%d.layout" % seed)
        with open(path, "w") as f:
            f.write(repr(d))
        horizon = int(rng.integers(20, 60))
        params = ns.get_base_layout_params(path, horizon)
        lp = layouts.load_layout(path, horizon)
        ref = run_reference(ns, path, lp, horizon, STEPS, rng, ScriptedTeam)
        k = "s%d_" % seed
        out[k + "layout"] = np.asarray(repr(d))
        out[k + "horizon"] = np.int32(horizon)
        out[k + "params"] = np.asarray(json.dumps(params, sort_keys=True))
        for name in ("actions", "rewards", "dones", "states", "reset_obs"):
            out[k + name] = ref[name]
        out[k + "obs_head"] = ref["obs"][:16]
        out[k + "obs_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(ref["obs"]).tobytes()).digest(), dtype=np.uint8)
        hist = {int(a): int(b) for a, b in zip(*np.unique(ref["rewards"], return_counts=True))}
        print(seed, "%dx%d P=%d horizon=%d" % (lp.width, lp.height, lp.num_players, horizon), hist)
    np.savez_compressed(os.path.join(HERE, "random_layouts.npz"), **out)
    print("wrote random_layouts.npz", os.path.getsize(os.path.join(HERE, "random_layouts.npz")), "bytes")


if __name__ == "__main__":
    main()
