#!/usr/bin/env python
"""Config-3 style exploration on the GPU box: builds the synthetic N-genome workload there (tools/synth_index), checks
a sample against the oracle, then times the device-resident path for several scan group widths, the C++ CLI end to end
and the reference CPU binary on the same index.
usage: perf_c3.py [--genomes 1000] [--length 3000000] [--reads 2000000] [--batch 1000000] [--groups 8,16,32] [--place]"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=1000)
    ap.add_argument("--length", type=int, default=3_000_000)
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--batch", type=int, default=250_000)
    ap.add_argument("--groups", default="staged")
    ap.add_argument("--dedup", action="store_true", help="count the distinct solves of one batch")
    ap.add_argument("--pipelines", default="sorted,fused", help="KREPP_PIPELINE values to time (sorted.cu / match.cu)")
    ap.add_argument("--out", default="/tmp/c3")
    ap.add_argument("--cpu-reads", type=int, default=100_000)
    ap.add_argument("--check", type=int, default=300)
    ap.add_argument("--place", action="store_true")
    ap.add_argument("--skip-cli", action="store_true")
    a = ap.parse_args()
    import torch
    import krepp_b200

    t0 = time.time()
    if not os.path.exists(os.path.join(a.out, "workload.json")):
        subprocess.run([os.path.join(ROOT, "tools", "_build", "synth_index"), "--out", a.out, "--genomes", str(a.genomes), "--length", str(a.length),
                        "--reads", str(a.reads), "--fastq-reads", str(a.reads)], check=True)
    wl = json.load(open(os.path.join(a.out, "workload.json")))
    print("workload", wl, f"built in {time.time() - t0:.1f} s", flush=True)
    idx = os.path.join(a.out, "index")
    reads = np.fromfile(os.path.join(a.out, "reads.u8"), dtype=np.uint8).reshape(-1, 150)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    if a.check:
        import oracle_lib as O
        from gpu_common import run_and_compare
        t = time.time()
        gi = krepp_b200.Index(idx, 0)
        print(f"index open {time.time() - t:.1f} s, device bytes {gi.info.device_bytes / 1e9:.2f} GB, mean bucket {gi.info.mean_bucket:.1f}", flush=True)
        sample = [r.tobytes() for r in reads[:a.check]]
        st = run_and_compare(idx, sample, O.OracleIndex(idx), gi, check_lookups=True)
        print("parity vs oracle (dist, all stages):", {k: st[k] for k in ("reads", "records", "solves", "bitexact_d", "max_rel_d")}, flush=True)
        st = run_and_compare(idx, sample, O.OracleIndex(idx), gi, check_lookups=False, place=True, no_filter=False)
        print("parity vs oracle (place):", {k: st[k] for k in ("reads", "placements")}, flush=True)
        gi.close()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    nb = (len(reads) + a.batch - 1) // a.batch
    d_reads = [torch.from_numpy(reads[i * a.batch:(i + 1) * a.batch].reshape(-1)).cuda() for i in range(nb)]
    for g in [(x, y) for y in a.pipelines.split(",") for x in a.groups.split(",")]:
        os.environ["KREPP_SCAN"], os.environ["KREPP_PIPELINE"] = g
        ix = krepp_b200.Index(idx, 0)
        b = krepp_b200.IBatch(ix, reads[:a.batch], place=a.place, no_filter=not a.place)
        d_o = torch.from_numpy(b.offsets.astype(np.int64)).cuda()
        res = []
        for it in range(3):
            mm = tt = 0.0
            alg = nrec = 0
            flush.zero_()
            torch.cuda.synchronize()
            w0 = time.time()
            for i in range(nb):
                n = len(d_reads[i]) // 150
                b.submit_device(d_reads[i].data_ptr(), d_o.data_ptr(), n, n * 150)
                r = b.wait()
                mm += r["match_ms"]; tt += r["gpu_ms"]; nrec += len(r["records"])
                alg += b.algorithmic_bytes()["bytes"]
            res.append((mm, tt, time.time() - w0))
        mm, tt, wall = min(res)
        ab = b.algorithmic_bytes()
        print("   stages of the last batch (ms):", "  ".join(f"{nm} {ms:.2f}" for nm, ms in b.stage_times()), flush=True)
        if a.dedup:  # how many distinct (leaf, histogram, mismatch count) solves the last batch holds
            rec, rd, hist = r["records"], r["reads"], r["hist"]
            solved = (rec["flags"] & 1) != 0
            key = np.column_stack([rec["leaf_se"][solved], hist[solved], rd["onmers"][rec["read"][solved]] - rec["match_count"][solved]]).astype(np.uint32)
            uniq = np.unique(key, axis=0)
            print(f"   solves in the last batch {int(solved.sum())}, distinct (leaf, hist, mismatch) {len(uniq)} -> x{solved.sum() / max(len(uniq), 1):.2f}; "
                  f"distinct (hist, mismatch) {len(np.unique(key[:, 1:], axis=0))}", flush=True)
            a.dedup = False
        print(f"scan={g} reads {len(reads)}  match {mm:8.2f} ms  kernels {tt:8.2f} ms  wall {wall * 1e3:8.1f} ms -> {len(reads) / tt / 1e3:7.2f} M reads/s  "
              f"algorithmic {alg / 1e9:7.1f} GB = {alg / len(reads) / 1e3:6.1f} kB/read -> {alg / mm / 1e6:7.1f} GB/s = {alg / mm / 1e6 / peak:5.3f} of measured HBM peak; "
              f"records/read {nrec / len(reads):5.1f}; entries/lookup {ab['entries'] / max(ab['lookups'], 1):5.1f}", flush=True)
        b.close(); ix.close()
    os.environ.pop("KREPP_SCAN", None)
    os.environ.pop("KREPP_PIPELINE", None)

    fq = os.path.join(a.out, "reads.fq")
    if not a.skip_cli:
        exe = os.path.join(ROOT, "krepp_b200", "_build", "krepp_b200")
        for threads in (1, os.cpu_count() or 1):
            p = subprocess.run([exe, "--num-threads", str(threads), "place" if a.place else "dist", "-i", idx, "-q", fq, "-o", "/dev/null"], capture_output=True, text=True)
            m = re.search(r"elapsed: ([0-9.eE+-]+) sec", p.stderr)
            print(f"CLI e2e ({threads} formatter threads): rc={p.returncode} {m.group(0) if m else p.stderr[-300:]} -> {len(reads) / float(m.group(1)) / 1e6 if m else 0:.2f} M reads/s", flush=True)
    if a.cpu_reads:
        sub = os.path.join(a.out, "cpu.fq")
        with open(fq, "rb") as f, open(sub, "wb") as g:
            for _ in range(4 * a.cpu_reads):
                g.write(f.readline())
        cores = os.cpu_count() or 1
        p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "krepp"), "--num-threads", str(cores), "place" if a.place else "dist", "-i", idx, "-q", sub, "-o", "/dev/null"],
                           capture_output=True, text=True)
        m = re.search(r"elapsed: ([0-9.eE+-]+) sec", p.stderr)
        print(f"reference CPU ({cores} threads, {a.cpu_reads} reads): {m.group(0) if m else p.stderr[-300:]} -> {a.cpu_reads / float(m.group(1)) / 1e3 if m else 0:.1f} k reads/s", flush=True)


if __name__ == "__main__":
    main()
