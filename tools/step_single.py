#!/usr/bin/env python
"""Device time of ONE env step per launch (ocb_step: the per-step path of policy rollouts and of n_step) at a given size.

    python tools/step_single.py --layout random1 --worlds 262144 --iters 20
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diverse_conventions_b200 import _native  # noqa: E402
from diverse_conventions_b200.overcooked_env import B200Overcooked, _ptr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layout", default="random1")
    ap.add_argument("--worlds", type=int, default=262144)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--lanes", type=int, default=0, help="lanes per world of the launch (0: the library's choice)")
    args = ap.parse_args()
    lib = _native.lib()
    env = B200Overcooked(args.layout, args.worlds, 0, horizon=400, seed=1)
    if args.lanes:
        env.set_tuning(args.lanes, True)
    N, P, SC = env.num_envs, env.num_players, env.width * env.height * (5 * env.num_players + 10)
    dev = env.sim_device
    obs = torch.empty((2, P, N, SC), dtype=torch.int8, device=dev)
    act = torch.randint(0, 6, (P, N), dtype=torch.int32, device=dev)
    rew = torch.empty((P, N), dtype=torch.int32, device=dev)
    done = torch.empty((N,), dtype=torch.int32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def step(i):
        _native.check(lib.ocb_step(env._h, _ptr(act), _ptr(obs[i & 1]), _ptr(rew), _ptr(done), stream))

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.iters):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / args.iters
    byts = P * N * SC + 8 * N * P + 4 * N
    print(json.dumps({"layout": args.layout, "worlds": N, "us_per_step": round(us, 2), "obs_gbs": round(byts / us / 1e3, 1),
                      "tuning": env.get_tuning()}))


if __name__ == "__main__":
    main()
