#!/usr/bin/env python
"""Where does one env step of the fused rollout kernel spend its time?

Runs ``ocb_rollout_fused_debug_trace`` (the instrumented build of ``rollout_fused_kernel``) and prints, for CTA 0 and a
few consecutive steps, the clock64 stamp of every hand-off of the dependent chain
env -> loader -> conv -> epilogue -> FC1 -> FC2 -> head -> sample -> env, relative to the moment the env warps received
the step's actions.  One JSON line per step plus a mean line.

    python tools/fused_trace.py --layout simple --worlds 8192 --T 40 --u0 8 --steps 8
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diverse_conventions_b200 import _native, layouts  # noqa: E402
from diverse_conventions_b200.overcooked_env import B200Overcooked, _ptr  # noqa: E402
from diverse_conventions_b200.policy import FusedPolicy, PolicyNet  # noqa: E402
from diverse_conventions_b200.rollout import RolloutBuffer  # noqa: E402

NAMES = {0: "env_got_actions", 1: "env_stepped", 2: "env_planes_free", 3: "env_planes_published", 8: "loader_sees_planes",
         9: "loader_first_col", 10: "loader_last_col", 48: "epi_sees_D2", 49: "epi_sees_D3", 50: "head_done",
         51: "actions_handed", 52: "critic_sees_D3", 53: "critic_head_done"}
NAMES.update({54: "conv3_issue_starts", 55: "FC2a_issue_starts", 56: "conv3_first_mma", 63: "conv3_mmas_issued", 57: "FC2a_critic_issue_starts",
              58: "FC2a_critic_issued", 59: "epi_conv2_in_regs", 60: "epi_conv2_split_done", 61: "epi_conv2_stored",
              62: "FC2a_operand_seen"})
NAMES.update({4: "loader_col1", 5: "loader_col2", 6: "loader_col3", 7: "loader_col4", 11: "conv0_issue_starts", 12: "conv1_issue_starts"})
for p in range(8):
    NAMES[16 + p] = "conv%d_issued" % p
    NAMES[24 + p] = "fc%d_issued" % p
    NAMES[32 + p] = "epi_sees_conv%d" % p
    NAMES[40 + p] = "epi_published%d" % p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layout", default="simple")
    ap.add_argument("--worlds", type=int, default=8192)
    ap.add_argument("--T", type=int, default=40)
    ap.add_argument("--u0", type=int, default=8)
    ap.add_argument("--steps", type=int, default=8)
    args = ap.parse_args()
    lp = layouts.load_layout(args.layout, 400)
    pol = FusedPolicy(lp, 64, 1)
    pol.set_weights(0, PolicyNet("actor", lp.width, lp.height, lp.channels, 64).init_like_reference(1),
                    PolicyNet("critic", lp.width, lp.height, lp.channels, 64).init_like_reference(2))
    env = B200Overcooked(args.layout, args.worlds, 0, horizon=400, seed=1)
    b = RolloutBuffer(env, args.T)
    trace = np.zeros((args.steps, 64), dtype=np.int64)
    lib = _native.lib()
    for _ in range(2):
        _native.check(lib.ocb_rollout_fused_debug_trace(env._h, pol._h, args.T, 0, _ptr(b.obs), _ptr(b.actions),
                                                        _ptr(b.action_log_probs), _ptr(b.value_preds), _ptr(b.rewards),
                                                        _ptr(b.dones), 1, trace.ctypes.data_as(ctypes.c_void_p), args.u0,
                                                        args.steps))
    torch.cuda.synchronize()
    rel_all = []
    for s in range(args.steps):
        t0 = trace[s, 0]
        row = {NAMES[e]: int(trace[s, e] - t0) for e in sorted(NAMES) if trace[s, e] != 0}
        if s + 1 < args.steps:
            row["next_env_got_actions"] = int(trace[s + 1, 0] - t0)
        rel_all.append(row)
        print(json.dumps({"step": args.u0 + s, **row}))
    keys = [k for k in rel_all[0] if all(k in r for r in rel_all[:-1])]
    mean = {k: round(float(np.mean([r[k] for r in rel_all[:-1]]))) for k in keys}
    print(json.dumps({"mean_cycles_since_env_got_actions": dict(sorted(mean.items(), key=lambda kv: kv[1])),
                      "layout": args.layout, "worlds": args.worlds}))


if __name__ == "__main__":
    main()
