#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass), condensed to the mnemonics that prove the
Blackwell paths: UTCHMMA (tcgen05.mma), UTCBAR (tcgen05.commit), LDTM / STTM (tcgen05.ld / st), UTCATOMSWS / UTC*ALLOC
(TMEM allocation), UBLKCP (cp.async.bulk), SYNCS (mbarrier), plus HMMA / FFMA / LDS / STS / LDG / STG totals.

    python tools/sass_histogram.py [libocb.so] > profiles/r2_sass_histogram.json
"""
import collections
import json
import os
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "SYNCS", "ELECT", "R2UR", "HMMA", "FFMA", "FMNMX", "PRMT",
        "LDS", "STS", "LDG", "STG", "REDG", "ATOMG", "SHFL", "BAR", "MUFU"]


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "diverse_conventions_b200", "libocb.so")
    text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|ocb::", "", name)
            name = re.sub(r"\(.*\)$", "", name).replace("void ", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    out = {"library": os.path.basename(so), "arch": "sm_100a", "kernels": {}}
    for name, c in kernels.items():
        row = {"instructions": c["_total"]}
        row.update({k: c[k] for k in KEYS if c[k]})
        out["kernels"][name] = row
    out["totals"] = {k: sum(c[k] for c in kernels.values()) for k in KEYS}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
