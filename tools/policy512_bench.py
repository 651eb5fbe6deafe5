#!/usr/bin/env python
"""Device-timed hidden-512 policy forward (conv512 + FC1 + FC2/head launches), CUDA events, one GPU."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diverse_conventions_b200 import layouts  # noqa: E402
from diverse_conventions_b200.overcooked_env import B200Overcooked  # noqa: E402
from diverse_conventions_b200.policy import FusedPolicy, PolicyNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layouts", default="simple,random1")
    ap.add_argument("--rows", type=int, default=32768)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    for layout in args.layouts.split(","):
        lp = layouts.load_layout(layout, 400)
        N = args.rows // 2
        env = B200Overcooked(layout, N, 0, horizon=400, seed=1)
        obs = env.rollout_random(37)["obs"][-1].contiguous()
        pol = FusedPolicy(lp, 512, 1)
        pol.set_weights(0, PolicyNet("actor", lp.width, lp.height, lp.channels, 512).init_like_reference(1),
                        PolicyNet("critic", lp.width, lp.height, lp.channels, 512).init_like_reference(2))
        out = pol.forward(obs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            pol.forward(obs, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        npos = (lp.width - 2) * (lp.height - 2)
        flops = 2 * args.rows * 2 * (180 * 256 * npos + 256 * npos * 512 + 512 * 512 + 512 * 6)
        print(json.dumps({"layout": layout, "hidden": 512, "rows": args.rows, "ms": round(ms, 4), "rows_per_s": round(args.rows / ms * 1e3),
                          "useful_TFLOPs": round(flops / ms / 1e9, 1), "issued_TFLOPs_bf16": round(3 * flops / ms / 1e9, 1)}), flush=True)
        env.close()
        pol.close()


if __name__ == "__main__":
    main()
