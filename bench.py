#!/usr/bin/env python
"""bench.py -- reads/s of the `krepp dist` hot path (BASELINE.json metric) on N B200s of one node.

Workload at every N (weak scaling: per-GPU work is fixed): BASELINE.json configs[1] -- the reference-built toy index
(25 genomes, -k 27 -w 35 -h 11) and 1,000,000 synthetic 150 bp reads per GPU, sampled from the toy genomes with
probability proportional to contig length, per-read substitution rate U(0, 0.15), random strand, numpy default_rng
seed 1 + rank (tools/synth.py).  A step = one pass of the hot path (match + solve + merge + finalize kernels) over
those reads.

  value        whole-job reads/s with the reads already resident in HBM (krepp_batch_submit_device), CUDA-event timed
  e2e          the same reads through the reference-facing C ABI with HOST buffers: krepp_batch_submit (H2D from
               pinned memory) + krepp_batch_wait (D2H of the result structs), 4 slots pipelined, wall-clock timed
  roofline     match kernel: algorithmic bytes (len + sum over lookups (16 + 8*|bucket|) + 64*records,
               counted on the device) / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the reference binary (oracle/_ref/krepp, unmodified, built by oracle/Makefile) on a bounded sample of
               the same reads with --num-threads = host cores

`--impl reference` times that reference CPU path alone and prints the same JSON shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

TOY = os.path.join(ROOT, "oracle", "_ref", "toy")
INDEX = os.path.join(TOY, "index_toy")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "krepp")
READS_PER_GPU = 1_000_000
READ_LEN = 150
METRIC = "reads/sec (krepp dist, 150bp)"
WORKLOAD = "configs[1]: toy index (25 genomes, k27 w35 h11), 1M synthetic 150bp reads per GPU, 0-15% substitutions"
# dram__bytes_read.sum + dram__bytes_write.sum of one match_kernel launch on this workload, from the ncu --set full
# capture summarised in profiles/ (None until a capture exists for the current kernel).
NCU_TRAFFIC_BYTES = None


def make_reads(n: int, seed: int) -> np.ndarray:
    import synth
    seq, offs = synth.load_packed(os.path.join(TOY, "genomes.npz"))
    return synth.sample_reads(seq, offs, n, read_len=READ_LEN, seed=seed)


def measured_peak_gbs() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 6:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(fastq: str, threads: int) -> tuple[float, int]:
    """Runs the unmodified reference CLI; returns (seconds of its own 'Done estimating distances' line, reads)."""
    p = subprocess.run([REF_BIN, "--num-threads", str(threads), "dist", "-i", INDEX, "-q", fastq, "-o", os.devnull],
                       capture_output=True, text=True, check=True)
    sec = float(re.search(r"Done estimating distances, elapsed: ([0-9.eE+-]+) sec", p.stderr).group(1))
    n = int(re.search(r"Total number of sequences queried: (\d+)", p.stderr).group(1))
    return sec, n


def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/krepp not built (needs /root/reference at build time)"}))
        return
    import synth
    cores = os.cpu_count() or 1
    sample = 100_000  # bounded sample of the 1M-read workload per step
    reads = make_reads(sample, seed=1)
    with tempfile.TemporaryDirectory() as td:
        fq = os.path.join(td, "sample.fq")
        synth.write_fastq(fq, reads)
        for _ in range(args.warmup):
            run_reference(fq, cores)
        secs = [run_reference(fq, cores)[0] for _ in range(args.steps)]
    t = sum(secs) / len(secs)
    v = sample / t
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample} of the 1M reads per step"},
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "reference",
                         "sample": f"{sample} reads per step, oracle/_ref/krepp --num-threads {cores} dist, its own elapsed line (index load excluded)"},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=READS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import krepp_b200

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the krepp_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert max(world, 1) == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    n = args.reads
    reads = make_reads(n, seed=1 + rank)
    index = krepp_b200.Index(INDEX, local)
    nbytes = n * READ_LEN

    # ---- device-resident arm (value): one slot, reads already in HBM
    slot = krepp_b200.IBatch(index, reads)
    d_bases = torch.from_numpy(slot.bases).cuda()
    d_offs = torch.from_numpy(slot.offsets.astype(np.int64)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        flush.zero_()  # evict the index and the reads from L2 between steps
        torch.cuda.synchronize()
        slot.submit_device(d_bases.data_ptr(), d_offs.data_ptr(), n, nbytes)
        return slot.wait()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    gpu_ms, match_ms, launches = [], [], 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = step_device()
        gpu_ms.append(r["gpu_ms"])
        match_ms.append(r["match_ms"])
        launches += r["gpu_launches"]
    barrier()
    wall_device = time.perf_counter() - t0
    clocks = sampler.stop()
    alg = slot.algorithmic_bytes()
    n_records = len(r["records"])
    t_dev = torch.tensor([sum(gpu_ms) / 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    t_dev = float(t_dev.item())
    value = world * n * args.steps / t_dev

    # ---- end-to-end arm (e2e): HOST buffers through krepp_batch_submit / krepp_batch_wait, 4 slots pipelined
    nslots, chunk = 4, (n + 3) // 4
    slots = []
    for c in range(nslots):
        sl = krepp_b200.IBatch(index, reads[c * chunk:(c + 1) * chunk])
        sl.pin_inputs()  # the step's inputs live in pinned host memory; every step copies them host->device again
        slots.append(sl)
    h2d = sum(int(s.offsets[-1]) + 8 * (s.n_reads + 1) for s in slots)

    def step_e2e():
        d2h = 0
        for s in slots:
            s.submit()
        for s in slots:
            res = s.wait()
            d2h += res["reads"].nbytes + res["records"].nbytes + res["hist"].nbytes
        return d2h

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d2h = step_e2e()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e = world * n * args.steps / float(t_e2e.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    mm = sum(match_ms) / len(match_ms)
    achieved = alg["bytes"] / (mm / 1e3) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32/f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reads_per_gpu": n, "read_len": READ_LEN, "seed": "numpy default_rng(1 + rank)",
                   "l2": "flushed between steps (256 MiB memset); the 72 MB toy index is re-read from HBM once per step and is L2-resident after that",
                   "index": "replicated per GPU", "records_per_step": n_records, "wall_s_device_arm": wall_device},
        "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "how": "krepp_batch_submit from pinned host buffers + krepp_batch_wait, 4 slots x 250k reads pipelined, wall clock"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "match_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC_BYTES, "algorithmic_bytes_per_launch": alg["bytes"], "lookups_per_launch": alg["lookups"],
                     "entries_scanned_per_launch": alg["entries"], "match_ms": mm, "match_share_of_step": mm / (t_dev * 1e3 / args.steps),
                     "peak_source": peak_src,
                     "note": "toy index is smaller than L2, so the algorithmic-byte rate measures gather throughput out of L2, not HBM (SURVEY.md 8d)"},
    }
    if not args.no_cpu_baseline and os.path.exists(REF_BIN):
        import synth
        cores = os.cpu_count() or 1
        sample = 200_000
        with tempfile.TemporaryDirectory() as td:
            fq = os.path.join(td, "sample.fq")
            synth.write_fastq(fq, reads[:sample])
            sec, nq = run_reference(fq, cores)
        out["cpu_baseline"] = {"value": nq / sec, "unit": "reads/s", "cores": cores, "kind": "reference",
                               "sample": f"first {sample} of the step's reads, oracle/_ref/krepp --num-threads {cores} dist -o /dev/null, "
                                         f"its own elapsed line ({sec:.2f} s, index load excluded)"}
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
