#!/usr/bin/env python
"""Where a self-play rollout step goes: CUDA-graph replays of (a) the fused policy forward alone,
(b) the single env step alone, (c) the full rollout, device-timed (config 4 shapes)."""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diverse_conventions_b200 import _native, layouts  # noqa: E402
from diverse_conventions_b200.overcooked_env import B200Overcooked, _ptr  # noqa: E402
from diverse_conventions_b200.policy import FusedPolicy, PolicyNet  # noqa: E402
from diverse_conventions_b200.rollout import PolicyRollout  # noqa: E402


def graph_time(fn, T, iters=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(T):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters / T


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layouts", default="simple,random1")
    ap.add_argument("--worlds", type=int, default=8192)
    ap.add_argument("--T", type=int, default=100)
    args = ap.parse_args()
    lib = _native.lib()
    for layout in args.layouts.split(","):
        lp = layouts.load_layout(layout, 400)
        N = args.worlds
        pol = FusedPolicy(lp, 64, 1)
        pol.set_weights(0, PolicyNet("actor", lp.width, lp.height, lp.channels, 64).init_like_reference(1),
                        PolicyNet("critic", lp.width, lp.height, lp.channels, 64).init_like_reference(2))
        env = B200Overcooked(layout, N, 0, horizon=400, seed=1)
        ro = PolicyRollout(env, pol, args.T, seed=1, use_graph=True)
        ro.collect()
        b = ro.buf
        obs0, obs1 = b.obs[0], b.obs[1]
        out = {"actions": b.actions[0].view(-1), "logp": b.action_log_probs[0].view(-1), "logits": None,
               "values": b.value_preds[0].view(-1)}
        stream = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        us_pol = graph_time(lambda: pol.forward(obs0, out=out), args.T)
        us_act = graph_time(lambda: pol.act(obs0, out=out), args.T)
        us_env = graph_time(lambda: _native.check(lib.ocb_step(env._h, _ptr(b.actions[0]), _ptr(obs1), _ptr(b.rewards[0]),
                                                               _ptr(b.dones[0]), stream())), args.T)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ro.collect()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            ro.collect()
        e1.record()
        torch.cuda.synchronize()
        us_full = 1e3 * e0.elapsed_time(e1) / 5 / args.T
        print(json.dumps({"layout": layout, "worlds": N, "us_policy_fused": round(us_pol, 2), "us_actor_only": round(us_act, 2),
                          "us_env_step": round(us_env, 2), "us_rollout_step": round(us_full, 2)}), flush=True)
        env.close()
        pol.close()


if __name__ == "__main__":
    main()
