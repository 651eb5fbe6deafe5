#!/usr/bin/env python
"""Device-timed returns / GAE pass over a config-4 sized rollout buffer (CUDA events, one GPU).
Algorithmic bytes per agent-step: value 4 + reward 4 + done 4/P (read), return 4 + advantage 4 (write)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diverse_conventions_b200 import returns as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=400)
    ap.add_argument("--worlds", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    T, P, N = args.T, 2, args.worlds
    v = torch.randn((T + 1, P, N), device="cuda")
    r = torch.randint(0, 2, (T, P, N), device="cuda", dtype=torch.int32)
    d = (torch.rand((T, N), device="cuda") < 0.01).to(torch.int32)
    ret = torch.zeros((T + 1, P, N), device="cuda")
    adv = torch.empty((T, P, N), device="cuda")
    for normalize in (False, True):
        for _ in range(3):
            R.compute_returns(v, r, d, normalize=normalize, out_returns=ret, out_advantages=adv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            R.compute_returns(v, r, d, normalize=normalize, out_returns=ret, out_advantages=adv)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        nbytes = T * P * N * 16 + T * N * 4 + (T * P * N * 8 if normalize else 0)
        print(json.dumps({"kernel": "gae" + ("+normalize" if normalize else ""), "T": T, "worlds": N, "ms": round(ms, 4),
                          "algorithmic_GBs": round(nbytes / ms / 1e6, 1),
                          "agent_steps_per_s": round(T * P * N / (ms * 1e-3))}), flush=True)


if __name__ == "__main__":
    main()
