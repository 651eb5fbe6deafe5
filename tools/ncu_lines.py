#!/usr/bin/env python
"""Per-source-line stall samples from an ncu report (needs -lineinfo and --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [--top 40] [--file policy_kernels.cu]
"""
import argparse
import csv
import io
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--file", default=None)
    args = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    cur_file, hdr, lines = None, None, []
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))
        try:
            samples = int(d["# Samples"])
        except Exception:
            continue
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
        lines.append((samples, cur_file, int(r[0]), r[1].strip(), stalls, d.get("Instructions Executed", "")))
    total = sum(x[0] for x in lines)
    print("total samples", total)
    sel = [x for x in lines if args.file is None or (x[1] or "").endswith(args.file)]
    for s, f, ln, src, st, ie in sorted(sel, key=lambda x: -x[0])[:args.top]:
        top = ", ".join("%s %d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print("%6d %5.1f%%  %s:%d  [%s] inst=%s\n         %s" % (s, 100.0 * s / max(total, 1), (f or "?").split("/")[-1], ln, top, ie, src[:120]))


if __name__ == "__main__":
    main()
