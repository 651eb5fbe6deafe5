#!/usr/bin/env python
"""Condense an `ncu --set full` report into a small JSON for profiles/:  python tools/ncu_summary.py rep.ncu-rep out.json
One entry per captured launch with the handful of counters the design notes cite."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_active.avg",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {}
    for i, h in enumerate(hdr):
        base = h.split(".", 2)[-1] if h.split(".")[0] in ("TPC", "SM_A", "SM_B", "SM_C") else h
        for k in KEEP:
            if h == k or base == k:
                col.setdefault(k, i)
    name_i = hdr.index("Kernel Name")
    launches = []
    for r in data:
        e = {"kernel": r[name_i]}
        for k, i in col.items():
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                v = r[i]
            e[k + (" [" + units[i] + "]" if units[i] else "")] = v
        launches.append(e)
    json.dump({"source": rep.split("/")[-1], "launches": launches}, open(out, "w"), indent=1)
    print(json.dumps(launches, indent=1))


if __name__ == "__main__":
    main()
