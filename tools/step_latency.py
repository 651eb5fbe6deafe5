#!/usr/bin/env python
"""Device time of the single-step API path (ocb_step, one launch per env step) per tuning."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diverse_conventions_b200.overcooked_env import B200Overcooked  # noqa: E402

for layout in ("simple", "unident_s"):
    for N in (1024, 16384, 65536):
        for G in (0, 1, 2, 4, 8):
            env = B200Overcooked(layout, N, 0, horizon=400)
            if G:
                env.set_tuning(G, True)
            a = torch.randint(0, 6, (2, N, 1), device="cuda", dtype=torch.int32)
            for _ in range(10):
                env.n_step(a)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                env.n_step(a)
            e1.record()
            torch.cuda.synchronize()
            print(json.dumps({"layout": layout, "N": N, "G": G or "default", "us_per_step": round(e0.elapsed_time(e1) * 1e3 / 200, 2),
                              "tuning": env.get_tuning()}), flush=True)
            env.close()
