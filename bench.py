#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Workload (BASELINE.json configs[2], the one the metric's target is quoted on):
Overcooked cramped_room, 2 agents, horizon 400, 16,384 worlds per GPU, uniform random
actions drawn on the device, env step + observation encode only.

A bench "step" is one pass of the hot path over one batch: ONE fused launch
(ocb_rollout_random) that advances every world of the job by `--env-steps-per-pass`
(default 100) environment steps and writes the [T, P, N, W, H, C] observation slab
(1.3 GB per GPU at the defaults, >> the 126 MB L2, so no flush is needed between
passes).  `value` = agent-steps (P x worlds x env steps) per second over all GPUs.
`e2e` runs the reference-facing single-step call with HOST buffers (ocb_step_host:
pinned actions H2D, kernel, obs + rewards + dones D2H, sync), T calls per pass.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAYOUT = "simple"  # cramped_room
HORIZON = 400
WORLDS_PER_GPU = 16384
METRIC = "agent-steps/sec (env+obs)"
UNIT = "agent-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000, help="timed passes (fused launches)")
    ap.add_argument("--warmup", type=int, default=30, help="untimed warm-up passes")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--env-steps-per-pass", type=int, default=100)
    ap.add_argument("--worlds", type=int, default=WORLDS_PER_GPU)
    ap.add_argument("--layout", default=LAYOUT)
    ap.add_argument("--lanes", type=int, default=0, help="lanes per world (kernel tuning), 0 = library default")
    ap.add_argument("--tma", type=int, default=-1, help="1/0 force the TMA bulk-store path, -1 = library default")
    ap.add_argument("--e2e-passes", type=int, default=3, help="passes of the host-buffer e2e measurement (<= --steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-policy-rollout", action="store_true", help="skip the '+policy fwd' leg (BASELINE config 4)")
    ap.add_argument("--policy-worlds", type=int, default=8192)
    ap.add_argument("--policy-T", type=int, default=100)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {"workload": "overcooked cramped_room (simple.layout), 2 agents, horizon %d, %d worlds/GPU, "
                        "random actions, env step + obs encode" % (HORIZON, args.worlds),
            "layout": args.layout, "worlds_per_gpu": args.worlds, "n_worlds": args.worlds * n_gpus,
            "horizon": HORIZON, "env_steps_per_pass": args.env_steps_per_pass,
            "step_definition": "1 bench step = 1 fused launch = %d env steps of every world" % args.env_steps_per_pass,
            "l2_policy": "outputs larger than L2 (obs slab per launch >> 126 MB); no flush needed",
            "parallelism": "worlds sharded, %d per GPU, no data-path collective" % args.worlds}


# ----------------------------------------------------------------------------- CPU baselines
def _py_port_worker(job):
    """one process: the Python port of the reference env on `n` worlds for `steps` steps"""
    layout, horizon, n, steps, seed = job
    import numpy as np
    from diverse_conventions_b200 import layouts
    from oracle.overcooked_oracle import OvercookedOracle
    lp = layouts.load_layout(layout, horizon)
    orc = OvercookedOracle(lp, n)
    acts = np.random.default_rng(seed).integers(0, 6, size=(steps, lp.num_players, n))
    t0 = time.perf_counter()
    for k in range(steps):
        orc.step(acts[k])
    return n * steps, time.perf_counter() - t0


def cpu_port_throughput(layout, steps, budget_s, cores):
    """agent-steps/s of the Python port (the reference's CPU path is per-world Python too,
    envs/overcooked2_reimplement.py + pantheonrl_extension/vectorenv.py:362-396), all cores."""
    import concurrent.futures as cf
    ws, dt = _py_port_worker((layout, HORIZON, 2, 50, 0))  # calibrate: seconds per world-step
    per_ws = dt / ws
    n_per_core = int(max(1, min(256, budget_s / (per_ws * steps))))
    jobs = [(layout, HORIZON, n_per_core, steps, 100 + i) for i in range(cores)]
    with cf.ProcessPoolExecutor(max_workers=cores) as ex:
        res = list(ex.map(_py_port_worker, jobs))
    total_ws = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return 2 * total_ws / wall, n_per_core * cores, steps


def c_port_throughput(layout, budget_s, threads):
    """agent-steps/s of the C restatement (oracle/ocb_oracle.c), `threads` host threads"""
    import concurrent.futures as cf
    import numpy as np
    from diverse_conventions_b200 import layouts
    from oracle.c_oracle import COracle
    lp = layouts.load_layout(layout, HORIZON)
    n, K = 2048, 50
    acts = np.random.default_rng(0).integers(0, 6, size=(K, 2, n)).astype(np.uint8)

    def work(_):
        orc = COracle(lp, n)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < budget_s:
            orc.rollout(acts)
            reps += 1
        return reps * n * K, time.perf_counter() - t0

    with cf.ThreadPoolExecutor(max_workers=threads) as ex:  # ctypes releases the GIL
        res = list(ex.map(work, range(threads)))
    return 2 * sum(r[0] for r in res) / max(r[1] for r in res)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bind_to_gpu_local_cores(gpu_index):
    """Pin this rank to the host cores NVML reports as local to its GPU, so that the pinned e2e buffers (first-touch
    placement) and the copy-issuing thread sit on the GPU's NUMA node.  Best effort: returns the core count used or
    None when NVML / the cpuset give nothing usable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        allowed = os.sched_getaffinity(0)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(allowed | {os.cpu_count() or 1}) + 64) // 64)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus = local & allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    c_oracle.build()
    cores = host_cores()
    t0 = time.perf_counter()
    # a pass of the CPU arm = env_steps_per_pass env steps over a bounded sample of worlds; the
    # number of passes is capped so that the whole run stays within ~20 s of CPU work
    env_steps = args.env_steps_per_pass * max(min(args.steps, 4), 1)
    # warm-up: a short run of the same worker (imports, allocator)
    _py_port_worker((args.layout, HORIZON, 2, max(min(args.warmup, 50), 3), 1))
    value, n_worlds, steps_done = cpu_port_throughput(args.layout, env_steps, 20.0, cores)
    c_value = c_port_throughput(args.layout, 3.0, cores)
    cfg = workload_config(args, args.gpus)
    sample = "%d worlds x %d steps over %d processes (python port of the reference env, 1 process per core)" % (
        n_worlds, steps_done, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * 2 * n_worlds * args.env_steps_per_pass / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "c_port_value": c_value,
                             "c_port_note": "oracle/ocb_oracle.c (C restatement, -O2), %d threads" % cores},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t_begin - 0.05 <= t <= t_end + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}



def ncu_traffic_bytes(kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel_name` from the newest committed
    `ncu --set full` summary under profiles/ (None if there is none for this kernel)"""
    import glob
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_full_*summary.json"))):
        try:
            with open(path) as f:
                rows = json.load(f)
            vals = []
            if isinstance(rows, dict):  # tools/ncu_summary.py: {"launches": [{"kernel": ..., "metric [unit]": value}]}
                for r in rows.get("launches", []):
                    if kernel_name.replace(" ", "") not in r.get("kernel", "").replace(" ", ""):
                        continue
                    tot = 0.0
                    for k, v in r.items():
                        if k.startswith("dram__bytes_read.sum [") or k.startswith("dram__bytes_write.sum ["):
                            tot += float(v) * unit[k[k.index("[") + 1:-1]]
                    vals.append(tot)
                rows = []
            for r in rows:
                if kernel_name.replace(" ", "") not in r.get("Kernel Name", "").replace(" ", ""):
                    continue
                tot = 0.0
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    num, u = r[key].split()
                    tot += float(num) * unit[u]
                vals.append(tot)
            if vals:
                best = statistics.mean(vals)
        except Exception:
            continue
    return best


def policy_rollout_leg(args, lp, local, rank, world, dev, barrier):
    import torch
    import torch.distributed as dist
    from diverse_conventions_b200.overcooked_env import B200Overcooked
    from diverse_conventions_b200.policy import FusedPolicy, PolicyNet
    from diverse_conventions_b200.rollout import PolicyRollout
    N, T = args.policy_worlds, args.policy_T
    pol = FusedPolicy(lp, 64, 1, gpu_id=local)
    pol.set_weights(0, PolicyNet("actor", lp.width, lp.height, lp.channels, 64).init_like_reference(1),
                    PolicyNet("critic", lp.width, lp.height, lp.channels, 64).init_like_reference(2))
    env = B200Overcooked(args.layout, N, local, horizon=HORIZON, seed=1, world_offset=rank * N)
    ro = PolicyRollout(env, pol, T, seed=1, use_graph=True)
    for _ in range(3):
        ro.collect()
    barrier()
    iters = 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ro.collect()
        ro.buf.compute_returns()  # GAE + advantage normalisation over the buffer just written
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_rollout = float(ms.item()) / iters
    out = {"metric": "agent-steps/sec (env+obs+policy fwd+buffer write+GAE)", "value": lp.num_players * N * world * T / (ms_rollout * 1e-3),
           "unit": UNIT, "worlds_per_gpu": N, "T": T, "hidden": 64, "ms_per_rollout": ms_rollout,
           "us_per_env_step": 1e3 * ms_rollout / T, "launches_per_rollout": 2 * T + 1 + 2,
           "what": "MAPPO self-play rollout: fused actor+critic tcgen05 forward of both seats, sampling, env step, "
                   "seat-major PPO buffer write, then returns/GAE; CUDA-graph replay, device-timed"}
    env.close()
    pol.close()
    return out

# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from diverse_conventions_b200 import layouts
    from diverse_conventions_b200.overcooked_env import B200Overcooked

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    numa_cores = bind_to_gpu_local_cores(local) if world > 1 else None

    N, K, W, spl = args.worlds, args.steps, args.warmup, args.env_steps_per_pass
    lp = layouts.load_layout(args.layout, HORIZON)
    P = lp.num_players
    env = B200Overcooked(args.layout, N, local, horizon=HORIZON, seed=0, world_offset=rank * N)
    if args.lanes or args.tma >= 0:
        env.set_tuning(args.lanes, bool(max(args.tma, 0)))
    out = env.alloc_rollout(spl, obs=True, actions=False)
    bytes_ws = layouts.io_bytes_per_world_step(lp)

    def run_passes(n_passes, events=None):
        for i in range(n_passes):
            timed = events is not None and (i % 16 == 0)  # sample per-launch durations without flooding events
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            env.rollout_random(spl, out)
            if timed:
                e1.record()
                events.append((e0, e1))
        return n_passes

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    run_passes(max(W, 3))
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t_begin = time.perf_counter()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    events = []
    start.record()
    launches = run_passes(K, events)
    stop.record()
    barrier()
    t_end = time.perf_counter()
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    ms = torch.tensor([start.elapsed_time(stop)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = P * N * world * K * spl / (ms_total * 1e-3)

    # average launch duration of the dominant kernel, CUDA events on the launching stream -> roofline
    launch_ms = statistics.mean(e0.elapsed_time(e1) for e0, e1 in events)
    k_launch = spl
    achieved = bytes_ws * N * k_launch / (launch_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs (sustained copy)"
    except Exception:
        pass

    # DRAM traffic per launch of the same kernel at the same shape, from the committed `ncu --set full` capture
    traffic = ncu_traffic_bytes("oc_rollout_kernel<%d, %d>" % (P, env.get_tuning()["lanes_per_world"])) \
        if (N, spl, args.layout) == (WORLDS_PER_GPU, 100, LAYOUT) else None

    # '+policy fwd' leg of the metric (BASELINE config 4 on this layout): MAPPO self-play rollout, random-init
    # actor + critic (hidden 64), on-device sampling, env step, PPO buffer write; one CUDA-graph replay per rollout
    policy_leg = None
    if not args.no_policy_rollout:
        policy_leg = policy_rollout_leg(args, lp, local, rank, world, dev, barrier)

    # end-to-end: the reference-facing single-step call with host buffers
    E = max(min(args.e2e_passes, K), 1) * spl  # single-step calls
    h_act = torch.randint(0, 6, (P, N), dtype=torch.int32).pin_memory()
    h_obs = torch.empty((P, N, lp.width, lp.height, lp.channels), dtype=torch.int8).pin_memory()
    h_rew = torch.empty((P, N), dtype=torch.int32).pin_memory()
    h_done = torch.empty((N,), dtype=torch.int32).pin_memory()
    for _ in range(5):
        env.step_host(h_act, h_obs, h_rew, h_done)
    barrier()
    t0 = time.perf_counter()
    for _ in range(E):
        env.step_host(h_act, h_obs, h_rew, h_done)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = P * N * world * E / float(e2e_s.item())
    h2d = 4 * P * N
    d2h = P * N * lp.size * lp.channels + 4 * P * N + 4 * N
    # same call without shipping the observation planes over PCIe (they normally stay on the device)
    t0 = time.perf_counter()
    for _ in range(E):
        env.step_host(h_act, None, h_rew, h_done)
    torch.cuda.synchronize()
    e2e_noobs = P * N * E / (time.perf_counter() - t0)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": workload_config(args, world),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": "oc_rollout_kernel<%d,%d>" % (P, env.get_tuning()["lanes_per_world"]),
                             "tuning": env.get_tuning(),
                             "bytes_per_world_step": bytes_ws, "world_steps_per_launch": N * k_launch,
                             "launch_ms": launch_ms, "peak_source": peak_src},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "call": "ocb_step_host (1 launch / step, obs+reward+done to pinned host memory)",
                        "calls": E, "value_without_obs_d2h_rank0": e2e_noobs, "gpu_local_cores_rank0": numa_cores},
                "policy_rollout": policy_leg,
                "gpu_launches": launches, "clocks": clocks,
                "env_steps": K * spl, "us_per_env_step": 1e3 * ms_total / (K * spl)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import c_oracle
                c_oracle.build()
                cores = host_cores()
                v, n_worlds, steps_done = cpu_port_throughput(args.layout, 400, args.cpu_seconds, cores)
                line["cpu_baseline"] = {
                    "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "%d worlds x %d steps, python port of the reference env, 1 process per core" % (
                        n_worlds, steps_done),
                    "c_port_value": c_port_throughput(args.layout, 2.0, cores),
                    "c_port_note": "oracle/ocb_oracle.c (C restatement), %d threads" % cores}
            except Exception as exc:  # the baseline must never take the GPU line down
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % exc}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
