#!/usr/bin/env python
"""Turns one gpurun_out/<tag>/ directory (written by tools/gpu_round.sh) into the tracked summary under profiles/:
   profiles/<tag>_launches.txt   per-kernel totals and shares from the ncu launch list (gpu__time_duration.sum)
   profiles/<tag>_<rep>.txt      selected raw metrics + hottest source lines of every *.ncu-rep capture (--set full)
   profiles/<tag>_bench.json     the bench lines of that run
usage: profile_summary.py gpurun_out/<tag> [out_tag [batch_reads_of_the_full_capture]]   (the third argument also writes
profiles/ncu_traffic.json from chain_full.ncu-rep)"""
import csv
import glob
import io
import os
import re
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
       "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
       "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "launch__occupancy_limit_warps", "sm__cycles_elapsed.max", "lts__t_sectors_srcunit_tex_op_read.sum",
       "lts__t_sectors_srcunit_tex_lookup_miss.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
       "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
       "smsp__thread_inst_executed_per_inst_executed.ratio"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) > iv:
            d[r[ik].split("(")[0]].append(float(r[iv].replace(",", "")))
    tot = sum(sum(x) for x in d.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare SHARES)\n# source: {path}\n")
        f.write(f"{'kernel':72s} {'launches':>8s} {'total_us':>12s} {'mean_us':>10s} {'share':>7s}\n")
        for k, x in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k[:72]:72s} {len(x):8d} {sum(x) / 1e3:12.1f} {sum(x) / len(x) / 1e3:10.1f} {sum(x) / tot:7.3f}\n")


def capture(rep, out, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; source: {rep}\n")
        if len(rows) >= 3:
            hdr, units = rows[0], rows[1]
            for li, vals in enumerate(rows[2:]):
                name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
                f.write(f"\n## launch {li}: {name[:100]}\n")
                for h, u, v in zip(hdr, units, vals):
                    if h in RAW:
                        f.write(f"{h:90s} {v:>18s} {u}\n")
        src = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, str(top)], capture_output=True, text=True).stdout
        f.write("\n## hottest source lines (share of executed warp instructions / of stall samples)\n" + src)


def traffic_json(rep, out, source, batch_reads):
    """profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel of the chain, as
    bench.py's roofline.traffic quotes them (valid for launches of `batch_reads` reads on the config-3 workload)."""
    import json
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    names = [("lookup_kernel<0", "lookup_kernel<count>"), ("lookup_kernel<1", "lookup_kernel<scatter>"), ("join_kernel", "join_kernel"),
             ("hit_scatter_kernel", "hit_scatter_kernel"), ("resolve_kernel", "resolve_kernel"), ("gate_kernel", "gate_kernel"),
             ("solve_kernel", "solve_kernel"), ("alias_kernel", "alias_kernel")]
    k = {}
    for vals in rows[2:]:
        for pat, key in names:
            if re.search(r"(^|[^a-z_])" + re.escape(pat), vals[ik]) and key not in k:
                k[key] = float(vals[ir].replace(",", "")) * mult[units[ir]] + float(vals[iw].replace(",", "")) * mult[units[iw]]
    with open(out, "w") as f:
        json.dump({"c3": {"source": source, "batch_reads": batch_reads, "kernels": k}}, f, indent=1)
        f.write("\n")


def main():
    d = sys.argv[1].rstrip("/")
    tag = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(d)
    if len(sys.argv) > 3 and os.path.exists(os.path.join(d, "chain_full.ncu-rep")):
        traffic_json(os.path.join(d, "chain_full.ncu-rep"), os.path.join(ROOT, "profiles", "ncu_traffic.json"), f"profiles/{tag}_chain_full.txt", int(sys.argv[3]))
    P = os.path.join(ROOT, "profiles")
    os.makedirs(P, exist_ok=True)
    for lc in glob.glob(os.path.join(d, "launches*.csv")):
        launches(lc, os.path.join(P, f"{tag}_{os.path.basename(lc)[:-4]}.txt"))
    for rep in glob.glob(os.path.join(d, "*.ncu-rep")):
        capture(rep, os.path.join(P, f"{tag}_{os.path.basename(rep)[:-8]}.txt"))
    lines = []
    for b in sorted(glob.glob(os.path.join(d, "bench*.json"))):
        lines += [l for l in open(b).read().splitlines() if l.strip()]
    if lines:
        with open(os.path.join(P, f"{tag}_bench.json"), "w") as f:
            f.write("\n".join(lines) + "\n")
    for extra in ("perf_match.log", "nvsmi.txt"):
        if os.path.exists(os.path.join(d, extra)):
            shutil.copy(os.path.join(d, extra), os.path.join(P, f"{tag}_{extra}"))


if __name__ == "__main__":
    main()
