#!/usr/bin/env python
"""Device-timed throughput of the fused policy forward alone (CUDA events, one GPU).

    python tools/policy_bench.py --layouts simple,random1 --rows 32768 --iters 50

Rows are real observations of random play.  Prints one JSON line per layout with the time per
forward, rows/s and the tensor-pipe floor fraction (MMA cycles the kernel must issue / elapsed)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diverse_conventions_b200 import layouts  # noqa: E402
from diverse_conventions_b200.overcooked_env import B200Overcooked  # noqa: E402
from diverse_conventions_b200.policy import FusedPolicy, PolicyNet  # noqa: E402


def mma_floor_cycles(npos, nets=2):
    """tcgen05 floor per 128-row tile (B300_MICROARCH: max(M,128)*N/256 cycles per K=16 MMA):
    conv 9 cells x (hi, lo) x N=32; FC1 2 k-steps x 3 products x N=64 per position; FC2 4 x 3 x N=64"""
    conv = npos * 18 * 16
    fc1 = npos * 6 * 32
    fc2 = 12 * 32
    return nets * (conv + fc1 + fc2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layouts", default="simple,random1")
    ap.add_argument("--rows", type=int, default=32768)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--mode", default="both", choices=["both", "act", "value", "fused"])
    args = ap.parse_args()
    sm_clock = 1.965e9
    for layout in args.layouts.split(","):
        lp = layouts.load_layout(layout, 400)
        N = args.rows // 2
        env = B200Overcooked(layout, N, 0, horizon=400, seed=1)
        out = env.rollout_random(37)
        obs = out["obs"][-1].contiguous()  # [P, N, W, H, C] == 2N rows
        try:
            pol = FusedPolicy(lp, 64, 1)
        except Exception as exc:
            print(json.dumps({"layout": layout, "error": str(exc)}), flush=True)
            continue
        pol.set_weights(0, PolicyNet("actor", lp.width, lp.height, lp.channels, 64).init_like_reference(1),
                        PolicyNet("critic", lp.width, lp.height, lp.channels, 64).init_like_reference(2))
        a = pol.act(obs)
        v = pol.value(obs)

        def run():
            if args.mode == "fused":
                pol.forward(obs)
            else:
                if args.mode in ("both", "act"):
                    pol.act(obs, out=a)
                if args.mode in ("both", "value"):
                    pol.value(obs, out=v)

        for _ in range(5):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        npos = (lp.width - 2) * (lp.height - 2)
        nets = 2 if args.mode in ("both", "fused") else 1
        tiles = (2 * N + 127) // 128
        floor_s = mma_floor_cycles(npos, nets) * -(-tiles // 148) / sm_clock
        print(json.dumps({"layout": layout, "rows": 2 * N, "mode": args.mode, "ms": round(ms, 4),
                          "rows_per_s": round(2 * N / (ms * 1e-3)), "mma_floor_ms": round(floor_s * 1e3, 4),
                          "tensor_floor_frac": round(floor_s * 1e3 / ms, 3)}), flush=True)
        env.close()
        pol.close()


if __name__ == "__main__":
    main()
