#!/usr/bin/env python
"""Kernel tuning sweep on one GPU: lanes-per-world x TMA x worlds (device-timed, CUDA events)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diverse_conventions_b200 import layouts  # noqa: E402
from diverse_conventions_b200.overcooked_env import B200Overcooked  # noqa: E402


def time_config(layout, N, G, tma, T, passes, obs=True):
    env = B200Overcooked(layout, N, 0, horizon=400, seed=0)
    env.set_tuning(G, bool(tma))
    out = env.alloc_rollout(T, obs=obs, actions=False)
    for _ in range(3):
        env.rollout_random(T, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(passes):
        env.rollout_random(T, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / passes
    env.close()
    del out
    torch.cuda.empty_cache()
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layouts", default="simple")
    ap.add_argument("--worlds", default="16384,65536,262144")
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--passes", type=int, default=20)
    ap.add_argument("--lanes", default="1,2,4,8,16", help="16 = role-split kernel (OCB_SPLIT_GE / OCB_SPLIT_TW pick its shape)")
    ap.add_argument("--quick", action="store_true", help="TMA + observations only")
    ap.add_argument("--tma", default="1,0")
    args = ap.parse_args()
    rows = []
    for layout in args.layouts.split(","):
        lp = layouts.load_layout(layout, 400)
        bws = layouts.io_bytes_per_world_step(lp)
        for N in [int(x) for x in args.worlds.split(",")]:
            T = args.T if N * lp.size * lp.channels * 2 * args.T < 8e9 else max(1, int(8e9 / (N * lp.size * lp.channels * 2)))
            for G in [int(x) for x in args.lanes.split(",")]:
                for tma in [int(x) for x in args.tma.split(",")]:
                    for obs in (True, False):
                        if not obs and not tma:
                            continue
                        if args.quick and not obs:
                            continue
                        ms = time_config(layout, N, G, tma, T, args.passes, obs)
                        gsteps = 2 * N * T / (ms * 1e-3) / 1e9
                        gbs = bws * N * T / (ms * 1e-3) / 1e9 if obs else 0.0
                        row = dict(layout=layout, N=N, T=T, G=G, tma=tma, obs=obs, ms=round(ms, 4),
                                   split_ge=os.environ.get("OCB_SPLIT_GE", ""), split_tw=os.environ.get("OCB_SPLIT_TW", ""),
                                   Gagent_steps_s=round(gsteps, 3), GBs=round(gbs, 1), frac=round(gbs / 6550.1, 3))
                        rows.append(row)
                        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
