#!/usr/bin/env python
"""Balance-Beam throughput on one B200 (BASELINE config 2: 2 agents, 65,536 worlds, random actions from the on-device
counter RNG): K fused steps per launch, device-timed, against the HBM roofline (76 algorithmic bytes per world-step:
2 x 7 int32 observations + 2 float rewards + 1 int32 done + 2 int32 actions).  One JSON line per setting."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diverse_conventions_b200.balance_env import B200BalanceBeam  # noqa: E402

BYTES_PER_WORLD_STEP = 76


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worlds", default="65536,1048576")
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--iters", type=int, default=100)
    args = ap.parse_args()
    peak = 6550.1
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for N in [int(x) for x in args.worlds.split(",")]:
        env = B200BalanceBeam(N, 0, seed=1)
        out = env.alloc_rollout(args.K, obs=True, actions=False)
        for _ in range(5):
            env.rollout_random(args.K, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            env.rollout_random(args.K, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        ws = N * args.K
        gbs = ws * BYTES_PER_WORLD_STEP / ms / 1e6
        print(json.dumps({"env": "balance_beam", "worlds": N, "steps_per_launch": args.K, "ms_per_launch": round(ms, 4),
                          "agent_steps_per_s": round(2 * ws / (ms * 1e-3)), "algorithmic_bytes_per_world_step": BYTES_PER_WORLD_STEP,
                          "GBps": round(gbs, 1), "hbm_peak_GBps": peak, "frac_of_hbm_peak": round(gbs / peak, 3),
                          "slab_MB": round(ws * 68 / 2**20, 1)}), flush=True)
        env.close()


if __name__ == "__main__":
    main()
