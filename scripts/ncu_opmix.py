#!/usr/bin/env python
"""Dynamic SASS opcode mix of a kernel from `ncu -i X.ncu-rep --page source --csv` (needs -lineinfo +
--import-source on).  usage: python scripts/ncu_opmix.py gpurun_out/prof.ncu-rep [n_warp_attempts]"""
import collections
import csv
import io
import re
import subprocess
import sys


def main(path, denom=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    c_src, c_ex, c_thr = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    c_stall = hdr.index("Warp Stall Sampling (All Samples)")
    mix = collections.Counter(); thr = collections.Counter(); stall = collections.Counter()
    total = 0
    for r in rows[hi + 1:]:
        if len(r) <= c_thr:
            continue
        m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_]+)", r[c_src])
        if not m:
            continue
        op = m.group(2)
        n = int(r[c_ex] or 0)
        mix[op] += n; thr[op] += int(r[c_thr] or 0); stall[op] += int(r[c_stall] or 0)
        total += n
    print(f"total warp instructions executed: {total}")
    if denom:
        print(f"per warp-attempt ({denom} warp-attempts): {total / denom:.0f}")
    tot_stall = sum(stall.values())
    for op, n in mix.most_common(32):
        line = f"  {op:10s} {n:>12d}  {100 * n / total:5.1f}%  stall-samples {100 * stall[op] / max(tot_stall, 1):5.1f}%"
        if denom:
            line += f"   {n / denom:7.1f}/attempt"
        print(line)


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
