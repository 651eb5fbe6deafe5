#!/usr/bin/env python
"""Instruction and stall-sample share per source line of one kernel in an ncu report (compiled with -lineinfo):

    python tools/ncu_lines_by_source.py report.ncu-rep path/to/kernels.cubin 'mangled kernel name substring' [top]

ncu's CSV source page lists SASS only; the line table comes from `nvdisasm -g` of the cubin
(`cuobjdump -xelf all libocb.so` extracts it) and is joined on the instruction offset."""
import csv
import io
import re
import subprocess
import sys
from collections import Counter


def main():
    rep, cubin, name = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and name in l)
    amap, cur = {}, None
    for l in dis[start + 1:]:
        if l.startswith(".text."):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", l)
        if m:
            amap[int(m.group(1), 16)] = cur
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[h], rows[h + 1:]
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(data[0][0], 16)
    c, cs = Counter(), Counter()
    for r in data:
        k = amap.get(int(r[0], 16) - base)
        c[k] += int(r[ie] or 0)
        cs[k] += int(r[isamp] or 0)
    tot, ts = sum(c.values()), sum(cs.values())
    print("warp instructions %d, stall samples %d" % (tot, ts))
    for k, v in c.most_common(top):
        print("%-28s inst %5.1f %%   samples %5.1f %%" % ("%s:%d" % k if k else "?", 100.0 * v / tot, 100.0 * cs[k] / max(ts, 1)))


if __name__ == "__main__":
    main()
