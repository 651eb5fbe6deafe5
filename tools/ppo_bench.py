#!/usr/bin/env python
"""PPO minibatch path on one B200: gather GB/s against the HBM roofline, in-place
evaluate_actions rows/s, loss launches.  Shape = BASELINE config 4: 8,192 worlds, T=100 self-play rollout,
2 seats -> 1,638,400 samples.  One JSON line per layout."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diverse_conventions_b200 import layouts, ppo  # noqa: E402
from diverse_conventions_b200.overcooked_env import B200Overcooked  # noqa: E402
from diverse_conventions_b200.policy import FusedPolicy, PolicyNet  # noqa: E402
from diverse_conventions_b200.rollout import PolicyRollout  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layouts", default="simple,random1,unident_s")
    ap.add_argument("--worlds", type=int, default=8192)
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--num-mini-batch", type=int, default=1)
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    peak = 6550.1
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for layout in args.layouts.split(","):
        lp = layouts.load_layout(layout, 400)
        pol = FusedPolicy(lp, args.hidden, 1)
        pol.set_weights(0, PolicyNet("actor", lp.width, lp.height, lp.channels, args.hidden).init_like_reference(1),
                        PolicyNet("critic", lp.width, lp.height, lp.channels, args.hidden).init_like_reference(2))
        env = B200Overcooked(layout, args.worlds, 0, horizon=400, seed=1)
        ro = PolicyRollout(env, pol, args.T, seed=1)
        buf = ro.collect()
        buf.compute_returns()
        rows = ppo.minibatch_rows(buf.T, buf.N, buf.P, args.num_mini_batch, torch.Generator().manual_seed(0), buf.obs.device)[0]
        B, SC = rows.numel(), lp.width * lp.height * lp.channels
        vn = ppo.ValueNormState("cuda")
        ev = pol.evaluate(buf.obs, buf.actions, rows)
        res = {"layout": layout, "hidden": args.hidden, "samples": B, "obs_bytes_per_row": SC}
        ms = timed(lambda: ppo.gather_minibatch(buf, rows, buf.advantages, buf.returns, torch.int8), args.iters)
        by = B * (2 * SC + 4 + 5 * 8)
        res["gather_int8"] = {"ms": round(ms, 4), "algorithmic_GB": round(by / 1e9, 4), "GBps": round(by / ms / 1e6, 1),
                              "frac_of_hbm_peak": round(by / ms / 1e6 / peak, 3)}
        ms = timed(lambda: ppo.gather_minibatch(buf, rows, buf.advantages, buf.returns, torch.float32), args.iters)
        by = B * (5 * SC + 4 + 5 * 8)
        res["gather_f32"] = {"ms": round(ms, 4), "algorithmic_GB": round(by / 1e9, 4), "GBps": round(by / ms / 1e6, 1),
                             "frac_of_hbm_peak": round(by / ms / 1e6 / peak, 3)}
        ms = timed(lambda: pol.evaluate(buf.obs, buf.actions, rows, out=ev), args.iters)
        res["evaluate_in_place"] = {"ms": round(ms, 4), "rows_per_s": round(B / ms * 1e3), "obs_GBps": round(B * SC / ms / 1e6, 1)}
        dense = torch.empty((B, lp.width, lp.height, lp.channels), dtype=torch.int8, device="cuda")
        dact = torch.zeros((B,), dtype=torch.int32, device="cuda")
        ms = timed(lambda: pol.evaluate(dense, dact, None, out=ev), args.iters)
        res["evaluate_dense"] = {"ms": round(ms, 4), "rows_per_s": round(B / ms * 1e3)}
        ms = timed(lambda: ppo.ppo_loss(rows, ev["logp"], ev["entropy"], ev["values"], buf.action_log_probs, buf.advantages,
                                        buf.value_preds, buf.returns, None, vn), args.iters)
        res["loss"] = {"ms": round(ms, 4), "rows_per_s": round(B / ms * 1e3)}
        # what the reference does for the same minibatch: fancy-index fp32 obs (+ scalars) out of [T*N*P, ...] tensors
        obs_f = buf.obs[:args.T].reshape(-1, lp.width, lp.height, lp.channels).float()
        idx = rows.long()
        ms = timed(lambda: (obs_f[idx], obs_f[idx]), max(2, args.iters // 3))
        res["torch_fp32_index_x2"] = {"ms": round(ms, 4), "note": "obs_batch + share_obs_batch fancy indexing of fp32 observations (shared_buffer.py:339-341)"}
        del obs_f
        print(json.dumps(res), flush=True)
        env.close()
        pol.close()


if __name__ == "__main__":
    main()
