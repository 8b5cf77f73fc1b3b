#!/usr/bin/env python
"""Ranks source lines of a kernel by executed instructions / stall samples from an .ncu-rep (needs -lineinfo).
usage: ncu_lines.py REPORT.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[h]
    iL, iS, iI, iT, iSm = hdr.index("Line No"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    num = lambda s: int(s) if s.isdigit() else 0
    lines, tot, tots = [], 0, 0
    for r in rows[h + 1:]:
        if len(r) > iSm and r[iL].isdigit():
            n, s = num(r[iI]), num(r[iSm])
            tot += n
            tots += s
            lines.append((n, s, num(r[iT]), int(r[iL]), r[iS][:120]))
    print("total warp-instructions", tot, "samples", tots)
    for n, s, t, l, src in sorted(lines, reverse=True)[:top]:
        print(f"{n / max(tot, 1) * 100:5.1f}% inst {s / max(tots, 1) * 100:5.1f}% smp  thr/inst {t / max(n, 1):5.1f}  L{l}: {src}")


if __name__ == "__main__":
    main()
