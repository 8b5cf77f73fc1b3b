#!/usr/bin/env python
"""Aggregates executed warp instructions and stall samples of a kernel by named source-line ranges.
usage: ncu_regions.py REPORT.ncu-rep name:lo-hi [name:lo-hi ...]"""
import csv, io, subprocess, sys


def main():
    rep = sys.argv[1]
    regions = []
    for a in sys.argv[2:]:
        n, r = a.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[h]
    iL, iI, iSm = hdr.index("Line No"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    agg = {n: [0, 0, [0] * len(stall_cols)] for n, _, _ in regions}
    agg["other"] = [0, 0, [0] * len(stall_cols)]
    num = lambda s: int(s) if s.isdigit() else 0
    tot = tots = 0
    for r in rows[h + 1:]:
        if len(r) > iSm and r[iL].isdigit():
            l = int(r[iL]); n = num(r[iI]); s = num(r[iSm])
            tot += n; tots += s
            key = next((nm for nm, lo, hi in regions if lo <= l <= hi), "other")
            agg[key][0] += n; agg[key][1] += s
            for j, c in enumerate(stall_cols):
                agg[key][2][j] += num(r[c])
    print(f"total warp-instructions {tot} samples {tots}")
    for k, (n, s, st) in agg.items():
        top = sorted(((v, hdr[stall_cols[j]]) for j, v in enumerate(st)), reverse=True)[:4]
        print(f"{k:14s} {n / max(tot,1) * 100:5.1f}% inst {s / max(tots,1) * 100:5.1f}% smp   " + "  ".join(f"{nm[6:]}={v / max(s,1) * 100:.0f}%" for v, nm in top))


if __name__ == "__main__":
    main()
