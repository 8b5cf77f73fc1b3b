#!/usr/bin/env python
"""bench.py -- reads/s of the `krepp dist` hot path (BASELINE.json metric) on N B200s of one node.

Default workload (every N; weak scaling: per-GPU work is fixed, index replicated, reads sharded, no data-path
collective): BASELINE.json configs[2] -- the configuration the metric and the HBM-gather roofline target are quoted on:
a synthetic 1,000-genome index (1,000 x 3 Mbp on a random binary tree, k27 w35 h11, ~1.6 GB of (k-mer, colour) entries,
~93 per bucket: far larger than L2) and 10,000,000 synthetic 150 bp reads PER GPU (0-15 % substitutions, random
strand).  The workload is generated on the box by tools/synth_index (tools/workload.py; cached under /tmp).
`--workload toy` runs configs[1] (reference-built toy index, 1 M reads per GPU) instead.

A step = one pass of the hot path (match + gate + solve + merge + finalize kernels) over the rank's reads, in batches.

  value        whole-job reads/s with the reads already resident in HBM (krepp_batch_submit_device); time = sum of the
               library's CUDA-event kernel spans of all batches of the step, max over ranks
  e2e          the same reads through the reference-facing C ABI with HOST buffers: krepp_batch_submit (H2D from
               page-locked host memory) + krepp_batch_wait (D2H of all result structs), 4 slots pipelined, wall clock
  roofline     match kernel: algorithmic bytes (len + sum over lookups (16 + 8*|bucket|) + 64*records,
               counted on the device) / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the reference binary (oracle/_ref/krepp, unmodified, built by oracle/Makefile) on a bounded sample of
               the same reads against the same index with --num-threads = host cores

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

import workload as W  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "krepp")
READ_LEN = W.READ_LEN
METRIC = "reads/sec (krepp dist, 150bp)"
DEFAULT_READS = {"c3": 10_000_000, "toy": 1_000_000, "c5": 4_000_000}
DEFAULT_BATCH = {"c3": 1_000_000, "toy": 1_000_000, "c5": 1_000_000}
WORKLOADS = {
    "c3": "configs[2]: synthetic 1,000-genome index (1,000 x 3 Mbp on a random binary tree, k27 w35 h11), 10M synthetic 150bp reads per GPU, "
          "0-15% substitutions, krepp dist, index replicated",
    "toy": "configs[1]: toy index (25 genomes, k27 w35 h11), 1M synthetic 150bp reads per GPU, 0-15% substitutions",
    "c5": "configs[4] analogue: index sharded by LSH bucket range over the GPUs (every GPU holds 1/N of the k-mer table; forced -- the "
          "1,000-genome table of configs[2] stands in for one that exceeds a GPU's HBM, a 10,000-genome index cannot be generated on the "
          "box within the bench's minutes), all-to-all of lookups and of hit entries over NCCL/NVLink, krepp dist",
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the same command
# at the same batch size (tools/gpu_round.sh -> profiles/<tag>_chain_full.txt), keyed by workload then kernel.
NCU_TRAFFIC = {"c3": {"source": "profiles/r06m_chain_full.txt", "batch_reads": 1_000_000, "kernels": {}}}
try:
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as _f:
        NCU_TRAFFIC = json.load(_f)
except Exception:
    pass


class Workload:
    def __init__(self, name: str, reads_per_gpu: int, rank: int, world: int, need_reads: bool = True):
        self.name = name
        if name == "toy":
            self.index = os.path.join(W.TOY, "index_toy")
            if not os.path.isdir(self.index):
                raise SystemExit("oracle/_ref/toy/index_toy is missing: run __graft_entry__.build() where /root/reference exists")
            self.reads = W.toy_reads(reads_per_gpu, seed=1 + rank) if need_reads else None
            self.info = {"seed": "numpy default_rng(1 + rank)"}
            self.fastq = None
            self.fastq_reads = 0
            self._seed = 1 + rank
        else:  # c3 and c5 share the generated index and read pool
            d, wl = W.ensure_c3(reads_per_gpu * world)
            self.index = os.path.join(d, "index")
            self.reads = W.c3_reads(d, rank * reads_per_gpu, reads_per_gpu) if need_reads else None
            self.info = {"generator": "tools/synth_index seed 7 (splitmix64), built on this box", "nkmers": wl["nkmers"],
                         "mean_bucket": wl["mean_bucket"], "nsubsets": wl["nsubsets"], "tree_nodes": wl["nnodes"],
                         "reads": "rank r takes reads [r*n, (r+1)*n) of the generated pool (a uniform sample)"}
            self.fastq = os.path.join(d, "reads.fq")  # first 200k reads of the pool = head of rank 0's reads
            self.fastq_reads = min(200_000, reads_per_gpu * world)
            self.dir = d


    def head(self, k: int) -> np.ndarray:
        """The first k reads of rank 0's reads."""
        if self.reads is not None and len(self.reads) >= k:
            return self.reads[:k]
        return W.toy_reads(k, seed=1) if self.name == "toy" else W.c3_reads(self.dir, 0, k)


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


def run_reference(index: str, fastq: str, threads: int, mode: str = "dist") -> tuple[float, int]:
    """Runs the unmodified reference CLI; returns (seconds of its own 'Done estimating distances' line, reads)."""
    p = subprocess.run([REF_BIN, "--num-threads", str(threads), mode, "-i", index, "-q", fastq, "-o", os.devnull],
                       capture_output=True, text=True, check=True)
    sec = float(re.search(r"Done (?:estimating distances|placing queries), elapsed: ([0-9.eE+-]+) sec", p.stderr).group(1))
    n = int(re.search(r"Total number of sequences queried: (\d+)", p.stderr).group(1))
    return sec, n


def sample_fastq(wl: Workload, td: str, sample: int) -> str:
    """FASTQ of the first `sample` reads of rank 0's reads."""
    fq = os.path.join(td, "sample.fq")
    if wl.fastq is not None and sample <= wl.fastq_reads:
        with open(wl.fastq, "rb") as f, open(fq, "wb") as g:
            for _ in range(4 * sample):
                g.write(f.readline())
    else:
        import synth
        synth.write_fastq(fq, wl.head(sample))
    return fq


def reference_sample(args, runs: int) -> int:
    """Reads per step of the reference arm: as many as keep the whole --steps/--warmup run within about four minutes (every run
    of the reference CLI also loads the index, ~4 s), between 50 k and 1 M.  The reference's rate still climbs slowly with
    the length of the run (71 k reads/s on 50 k reads, 78 k on 1 M), so longer is fairer to it."""
    if args.cpu_sample:
        return args.cpu_sample
    if args.workload == "toy":
        return 100_000
    per_run_s = max(1.0, 230.0 / max(runs, 1) - 4.0)
    return int(min(1_000_000, max(50_000, 75_000 * per_run_s)))


def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/krepp not built (needs /root/reference at build time)"}))
        return
    cores = os.cpu_count() or 1
    sample = reference_sample(args, args.steps + args.warmup)
    n = args.reads or DEFAULT_READS[args.workload]
    sample = min(sample, n)
    wl = Workload(args.workload, n, 0, max(args.gpus, 1), need_reads=False)  # the same pool (and cache directory) as the B200 arm's
    with tempfile.TemporaryDirectory() as td:
        fq = sample_fastq(wl, td, sample)
        for _ in range(args.warmup):
            run_reference(wl.index, fq, cores, args.mode)
        secs = [run_reference(wl.index, fq, cores, args.mode)[0] for _ in range(args.steps)]
    t = sum(secs) / len(secs)
    v = sample / t
    print(json.dumps({
        "impl": "reference", "metric": METRIC if args.mode == "dist" else METRIC.replace("dist", "place"), "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "sample": f"{sample} of the step's reads per step", **wl.info},
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "reference",
                         "sample": f"{sample} reads per step, oracle/_ref/krepp --num-threads {cores} {args.mode} -o /dev/null, its own elapsed line (index load excluded)"},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


class Env:
    """torch / torch.distributed plumbing shared by the arms: one process per GPU, NCCL for the barrier and the max over ranks."""

    def __init__(self, gpus: int):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the krepp_b200 hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        assert max(self.world, 1) == gpus or self.world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def cpu_baseline(wl: Workload, mode: str, sample: int) -> dict | None:
    if not os.path.exists(REF_BIN):
        return None
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as td:
        fq = sample_fastq(wl, td, sample)
        sec, nq = run_reference(wl.index, fq, cores, mode)
    return {"value": nq / sec, "unit": "reads/s", "cores": cores, "kind": "reference",
            "sample": f"first {sample} of rank 0's reads, oracle/_ref/krepp --num-threads {cores} {mode} -o /dev/null on the same index, "
                      f"its own elapsed line ({sec:.2f} s, index load excluded)"}


def instruction_roof(entries_per_launch: float, ms_per_launch: float, sm_mhz: float | None, sms: int = 148) -> dict:
    """Lower bounds on join_kernel from its instruction mix: one comparison of an index entry with a query is LOP3, LOP3, POPC,
    ISETP (XOR of the two bit-planes, OR, population count, threshold).  POPC issues at 16 lanes per clock per SM (quarter rate,
    CUDA C Programming Guide arithmetic-throughput table, cc 10.0), the other three at 64; four schedulers issue one warp
    instruction per clock each."""
    clk = (sm_mhz or 1965.0) * 1e6
    popc_ms = entries_per_launch / (sms * 16.0 * clk) * 1e3
    issue_ms = 4.0 * entries_per_launch / (sms * 128.0 * clk) * 1e3
    roof = max(popc_ms, issue_ms)
    return {"comparisons_per_launch": entries_per_launch, "popc_pipe_ms": popc_ms, "issue_ms": issue_ms, "bound": "popc pipe" if popc_ms >= issue_ms else "issue slots",
            "roof_ms": roof, "frac": roof / ms_per_launch if ms_per_launch else None,
            "assumes": f"{sms} SMs at {clk / 1e6:.0f} MHz, POPC 16 lanes/clk/SM, 4 instructions per comparison"}


def hot_arm(env: Env, index, wl: Workload, mode: str, n: int, batch: int, steps: int, warmup: int, e2e_batch: int, want_e2e: bool,
            cpu_sample: int, workload_name: str, t_wl: float) -> dict | None:
    """One arm of the replicated-index path (mode A): `dist` (configs[2]) or `place` (configs[3]).  Runs on every rank; rank 0
    gets the result object."""
    import krepp_b200
    torch = env.torch
    world = env.world
    reads = wl.reads[:n]
    mode_kw = dict(place=True, no_filter=False) if mode == "place" else {}

    def front_end_rows(s):  # what the command line asks a slot for (cli.cpp)
        if mode == "dist":
            s.set_output(records=False, hist=False, placements=False, summaries=False, dist=True)
        else:
            s.set_output(records=False, hist=False)  # placement rows + 44-byte read summaries

    # ---- device-resident arm (value): one slot, all reads of the step already in HBM, processed in batches
    slot = krepp_b200.IBatch(index, reads[:batch], **mode_kw)
    front_end_rows(slot)
    d_bases = torch.from_numpy(reads.reshape(-1)).cuda()
    d_offs = torch.from_numpy(slot.offsets.astype(np.int64)).cuda()  # fixed-length reads: every batch has the same offsets
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    chunks = [(i, min(batch, n - i)) for i in range(0, n, batch)]

    def step_device():
        flush.zero_()  # evict the index and the reads from L2 between steps
        torch.cuda.synchronize()
        gpu = mm = 0.0
        launches = nrec = alg = lk = en = npl = 0
        stages: dict = {}
        for first, cnt in chunks:
            slot.submit_device(d_bases.data_ptr() + first * READ_LEN, d_offs.data_ptr(), cnt, cnt * READ_LEN)
            r = slot.wait_device()  # rows stay in HBM; the e2e arm below is the one that copies them out
            gpu += r["gpu_ms"]; mm += r["match_ms"]; launches += r["gpu_launches"]; nrec += r["n_records"]; npl += r["n_placements"]
            ab = slot.algorithmic_bytes()
            alg += ab["bytes"]; lk += ab["lookups"]; en += ab["entries"]
            for name, ms in slot.stage_times():
                stages[name] = stages.get(name, 0.0) + ms
        return dict(gpu_ms=gpu, match_ms=mm, launches=launches, records=nrec, alg=alg, lookups=lk, entries=en, stages=stages, placements=npl)

    for _ in range(warmup):
        step_device()
    env.barrier()
    sampler = ClockSampler(env.local)
    sampler.start()
    gpu_ms, match_ms, launches = [], [], 0
    stage_ms: dict = {}
    t0 = time.perf_counter()
    for _ in range(steps):
        r = step_device()
        gpu_ms.append(r["gpu_ms"])
        match_ms.append(r["match_ms"])
        launches += r["launches"]
        for name, ms in r["stages"].items():
            stage_ms[name] = stage_ms.get(name, 0.0) + ms / steps
    env.barrier()
    wall_device = time.perf_counter() - t0
    clocks = sampler.stop()
    last = r
    t_dev = env.max(sum(gpu_ms) / 1e3)
    value = world * n * steps / t_dev
    slot.close()
    del d_bases

    # ---- end-to-end arm (e2e): HOST buffers through krepp_batch_submit / krepp_batch_wait, 4 slots pipelined
    e2e_out = None
    if want_e2e:
        eb = min(e2e_batch, n)
        nslots = 4
        h_reads = torch.from_numpy(reads.reshape(-1)).pin_memory()  # the step's inputs live in page-locked host memory
        h_offs = (np.arange(eb + 1, dtype=np.uint64) * np.uint64(READ_LEN))
        slots = [krepp_b200.IBatch(index, reads[:eb], **mode_kw) for _ in range(nslots)]
        for s_ in slots:
            front_end_rows(s_)
        echunks = [(i, min(eb, n - i)) for i in range(0, n, eb)]
        h2d = n * READ_LEN + 8 * sum(c + 1 for _, c in echunks)
        base_ptr = h_reads.data_ptr()
        row_keys = ("reads", "records", "hist", "placements", "brief", "dist_begin", "dist_rows")

        def step_e2e():
            d2h, nrec, nrow, inflight = 0, 0, 0, []

            def take(s):
                nonlocal d2h, nrec, nrow
                res = s.wait()
                d2h += sum(res[k].nbytes for k in row_keys)
                nrec += res["n_records"]
                nrow += len(res["dist_rows"]) + len(res["placements"])

            for j, (first, cnt) in enumerate(echunks):
                s = slots[j % nslots]
                if len(inflight) == nslots:
                    take(inflight.pop(0))
                s.submit_host(base_ptr + first * READ_LEN, h_offs.ctypes.data, cnt)
                inflight.append(s)
            for s in inflight:
                take(s)
            return d2h, nrec, nrow

        for _ in range(warmup):
            step_e2e()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            d2h, nrec_e2e, nrow_e2e = step_e2e()
        env.barrier()
        t_e2e = env.max(time.perf_counter() - t0)
        e2e = world * n * steps / t_e2e
        rows_how = ("4 bytes per read (row offsets + the NA flag) and 4 bytes per printed TSV row (reference << 16 | distance as the integer its five printed decimals show): "
                    "the rows `krepp dist` prints are selected, ordered and rounded by a kernel (KREPP_OUT_DIST), nothing else leaves the device"
                    if mode == "dist" else "the 44-byte read summaries and the 56-byte placement rows (all the jplace writer reads); records and Hamming histograms stay in HBM")
        e2e_out = {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "records_per_step": nrec_e2e, "output_rows_per_step": nrow_e2e,
                   "d2h_bytes_per_read": d2h / n, "h2d_bytes_per_read": h2d / n,
                   "how": f"krepp_batch_submit from page-locked host memory + krepp_batch_wait with the output the command line asks for (krepp_batch_set_output): {rows_how}; "
                          f"{nslots} slots x {eb} reads pipelined, wall clock, max over ranks"}
        for s in slots:
            s.close()
        del h_reads

    if env.rank != 0:
        return None

    peak, peak_src = measured_peak_gbs()
    mm = sum(match_ms) / len(match_ms)            # per step
    launches_per_step = len(chunks)
    achieved = last["alg"] / (mm / 1e3) / 1e9
    # The match step (ref src/query.cpp:40-94,352-390) is ONE kernel in the fused pipeline and a short chain of kernels in
    # the bucket-sorted one; the roofline is quoted over the whole chain (every launch between the library's two match
    # events), which is the conservative reading: the algorithmic bytes are those of the step, the time is all of it.
    sorted_pipeline = "join_kernel" in stage_ms
    match_name = ("match step, bucket-sorted pipeline: lookup_kernel x2 + scans + join_kernel + hit_scatter_kernel + resolve_kernel"
                  if sorted_pipeline else "match_kernel")
    dom = max(stage_ms, key=stage_ms.get) if stage_ms else "match_kernel"
    # ncu DRAM traffic of the match step's kernels, per launch of `batch` reads (null when the capture was taken at another batch size)
    nt = NCU_TRAFFIC.get("c3" if wl.name != "toy" else "toy", {})
    kt = nt.get("kernels", {}) if nt.get("batch_reads") == batch and mode == "dist" else {}
    chain = ["lookup_kernel<count>", "lookup_kernel<scatter>", "join_kernel", "hit_scatter_kernel", "resolve_kernel"] if sorted_pipeline else ["match_kernel"]
    traffic = sum(kt[k] for k in chain) if kt and all(k in kt for k in chain) else None
    # the dominant kernel on its own.  For join_kernel the algorithmic bytes are the bucket bytes of every lookup (8 B per entry
    # scanned, SURVEY 8d: no credit for reuse); the kernel reads each bucket ONCE per batch for all its lookups, which is why its
    # algorithmic rate exceeds the HBM peak while its DRAM traffic is a small fraction of the algorithmic bytes: its physical
    # roof is the instruction roof quoted beside it.
    dom_obj = {"name": dom, "ms_per_launch": stage_ms.get(dom, 0.0) / launches_per_step,
               "share_of_step": stage_ms.get(dom, 0.0) / (t_dev * 1e3 / steps)}
    join_obj = None
    if sorted_pipeline:
        jms = stage_ms.get("join_kernel", 0.0) / launches_per_step
        ab = 8.0 * last["entries"] / launches_per_step
        join_obj = {"name": "join_kernel", "ms_per_launch": jms, "algorithmic_bytes_per_launch": ab, "achieved": ab / (jms / 1e3) / 1e9 if jms else None, "unit": "GB/s",
                    "frac": ab / (jms / 1e3) / 1e9 / peak if jms else None, "traffic": kt.get("join_kernel"),
                    "instr_roof": instruction_roof(last["entries"] / launches_per_step, jms, clocks.get("sm_mhz"))}
        if dom == "join_kernel":
            dom_obj.update(join_obj)
    out = {
        "metric": METRIC if mode == "dist" else METRIC.replace("dist", "place"), "value": value, "unit": "reads/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": t_dev * 1e3 / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32/f64", "data": "synthetic",
        "config": {"workload": workload_name, "reads_per_gpu": n, "read_len": READ_LEN, "batch_reads": batch, **wl.info,
                   "l2": "inputs larger than L2 (index image %.2f GB, reads %.2f GB per step) and a 256 MiB memset between steps" % (index.info.device_bytes / 1e9, n * READ_LEN / 1e9),
                   "index": "replicated per GPU", "records_per_step": last["records"], "placements_per_step": last["placements"], "wall_s_device_arm": wall_device,
                   "workload_setup_s": round(t_wl, 1)},
        "e2e": e2e_out,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": match_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "stages_ms_per_step": {k: round(v, 3) for k, v in stage_ms.items()},
                     "dominant_kernel": dom_obj, "join_kernel": join_obj,
                     "whole_step_frac": (last["alg"] / (t_dev / steps) / 1e9) / peak,
                     "traffic": traffic, "traffic_source": nt.get("source"),
                     "algorithmic_bytes_per_launch": last["alg"] / launches_per_step,
                     "lookups_per_launch": last["lookups"] / launches_per_step, "entries_scanned_per_launch": last["entries"] / launches_per_step,
                     "launches_per_step": launches_per_step, "match_ms_per_launch": mm / launches_per_step,
                     "match_share_of_step": mm / (t_dev * 1e3 / steps), "peak_source": peak_src,
                     "roofline_reads_per_s": peak * 1e9 / (last["alg"] / n)},
    }
    out["cpu_baseline"] = cpu_baseline(wl, mode, min(cpu_sample, n)) if cpu_sample else None
    return out


def shard_arm(env: Env, wl: Workload, n: int, batch: int, steps: int, warmup: int, want_e2e: bool, t_wl: float, cpu_sample: int, lanes: int = 2) -> dict | None:
    """Mode B (BASELINE configs[4]): the index does not fit the per-GPU memory budget, so every rank holds one
    bucket-range shard of the table and its own reads; a step = every rank's reads through lookup -> all-to-all -> join on the
    owning shard -> all-to-all -> resolve / solve, batch by batch.  The budget is set so that the index needs exactly `world`
    shards (krepp_index_plan_shards).  Timed with CUDA events on torch's stream around the step (the library's calls return only
    once their kernels are done, and the NCCL exchanges run on torch's stream), max over ranks."""
    import krepp_b200.dist as kd
    from krepp_b200 import capi
    torch, dist = env.torch, env.dist
    rank, world, local = env.rank, env.world, env.local
    reads = wl.reads[:n]
    whole = capi.plan_shards(wl.index, 1 << 62, local)["whole_bytes"]
    table = 8 * wl.info["nkmers"] if "nkmers" in wl.info else whole
    budget = int(whole - table + table / world * 1.03) if world > 1 else whole
    plan = capi.plan_shards(wl.index, budget, local)
    assert plan["nshards"] == world, (plan, world, budget)
    me = kd.ShardRank(wl.index, local, rank, world, batch, batch * READ_LEN + 64, lanes=lanes)  # two batches in flight: one's tail under the next one's head
    job = kd.ShardedJob([me])
    h_reads = torch.from_numpy(reads.reshape(-1)).pin_memory()
    d_bases = torch.empty(n * READ_LEN + 64, dtype=torch.uint8, device="cuda")
    d_bases[:n * READ_LEN].copy_(h_reads)
    d_offs = (torch.arange(batch + 1, dtype=torch.int64, device="cuda") * READ_LEN)
    stage = [torch.empty(batch * READ_LEN + 64, dtype=torch.uint8, device="cuda") for _ in me.lanes]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    chunks = [(i, min(batch, n - i)) for i in range(0, n, batch)]

    def step(from_host: bool):
        tot = dict(nrec=0, d2h=0)
        alg = dict(bytes=0, lookups=0, entries=0)

        def batches():
            for j, (first, cnt) in enumerate(chunks):
                if from_host:
                    src = stage[j % len(stage)]
                    src[:cnt * READ_LEN].copy_(h_reads[first * READ_LEN:(first + cnt) * READ_LEN], non_blocking=True)
                else:
                    src = d_bases[first * READ_LEN:]
                yield [(src[:cnt * READ_LEN + 64], d_offs, cnt)]

        def consume(i, res, lane):
            r = res[0]
            tot["nrec"] += r["n_records"]
            if from_host:
                tot["d2h"] += r["dist_begin"].nbytes + r["dist_rows"].nbytes
            ab = me.lanes[lane].slot.algorithmic_bytes()
            for k in alg:
                alg[k] += ab[k]

        job.run_stream(batches(), rows=from_host, consume=consume)
        return tot["nrec"], tot["d2h"], alg

    def timed(from_host: bool):
        for _ in range(warmup):
            step(from_host)
        env.barrier()
        job.bytes_exchanged = 0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            if not from_host:
                flush.zero_()
            out = step(from_host)
        ev1.record()
        env.barrier()
        wall = time.perf_counter() - t0
        return env.max(ev0.elapsed_time(ev1) / 1e3), env.max(wall), out, job.bytes_exchanged / steps

    sampler = ClockSampler(local)
    sampler.start()
    for ln in me.lanes:
        ln.slot.set_output(records=False, hist=False, placements=False, summaries=False, dist=True)  # what the dist front end asks for
    t_dev, _, (nrec, _, alg), xbytes = timed(False)
    clocks = sampler.stop()
    stages = dict(me.slot.stage_times())
    t_e2e, d2h = 0.0, 0
    if want_e2e:
        _, t_e2e, (_, d2h, _), _ = timed(True)
    tot = kd.sum_over_ranks([alg["bytes"], alg["lookups"], alg["entries"], nrec, int(xbytes)], device="cuda")
    sh = me.index.shard
    image = me.index.info.device_bytes
    out = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = tot[0] / (t_dev / steps) / 1e9
        xch_ms = sum(v for k, v in stages.items() if k.startswith("(exchange"))
        out = {
            "metric": METRIC, "value": world * n * steps / t_dev, "unit": "reads/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": t_dev * 1e3 / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/f64", "data": "synthetic",
            "config": {"workload": WORKLOADS["c5"], "reads_per_gpu": n, "read_len": READ_LEN, "batch_reads": batch, "batches_in_flight": lanes, **wl.info,
                       "budget": f"per-GPU image budget {budget / 1e9:.3f} GB: the whole image is {whole / 1e9:.3f} GB and needs {plan['nshards']} bucket-range shards to fit "
                                 f"(krepp_index_plan_shards); shard image {image / 1e9:.3f} GB",
                       "index": f"sharded by bucket range, {world} shards; rank 0 holds rows [{sh.row0}, {sh.row1}) = {sh.n_entries} of {me.index.info.nkmers} entries",
                       "l2": "inputs larger than L2 and a 256 MiB memset between steps", "records_per_step": tot[3], "workload_setup_s": round(t_wl, 1)},
            "e2e": {"value": world * n * steps / t_e2e, "unit": "reads/s", "h2d_bytes_per_step": n * READ_LEN, "d2h_bytes_per_step": d2h,
                    "how": "every batch copied from page-locked host memory inside the timed region, the printed rows (KREPP_OUT_DIST) copied back by krepp_batch_wait; wall clock, max over ranks"} if want_e2e else None,
            "gpu_launches": steps * len(chunks) * (16 + world), "clocks": clocks,  # per batch: 5 lookup/scan + one join per sender + 6 regroup/resolve + 5 gate..dist rows
            "exchange": {"bytes_received_per_step_all_ranks": tot[4], "per_read": tot[4] / (world * n),
                         "transport": "torch.distributed all_to_all_single (NCCL over NVLink / NVSwitch)" if world > 1 else "none (one shard)",
                         "last_batch_exchange_ms_rank0": round(xch_ms, 3),
                         "nvlink_gbs_per_gpu_during_exchange": (tot[4] / world / len(chunks)) / (xch_ms / 1e3) / 1e9 if xch_ms else None},
            "roofline": {"bound": "hbm", "kernel": "whole step of all ranks (lookup, exchange, join on the owning shard, exchange, resolve, solve)", "achieved": achieved,
                         "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world), "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                         "algorithmic_bytes_per_step": tot[0], "lookups_per_step": tot[1], "entries_scanned_per_step": tot[2],
                         "last_batch_stages_ms_rank0": {k: round(v, 3) for k, v in stages.items()}},
        }
        out["cpu_baseline"] = cpu_baseline(wl, "dist", min(cpu_sample, n)) if cpu_sample else None
    me.close()
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "toy", "c5"])
    ap.add_argument("--mode", default="dist", choices=["dist", "place"], help="place = BASELINE configs[3] as the main line: krepp place (placement kernels on top of dist)")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (default: 10M for c3, 1M for toy)")
    ap.add_argument("--batch", type=int, default=0, help="reads per batch of the device-resident arm")
    ap.add_argument("--e2e-batch", type=int, default=500_000)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-place", action="store_true", help="leave out the `place` object (configs[3]) of the default line")
    ap.add_argument("--no-mode-b", action="store_true", help="leave out the `mode_b` object (configs[4], N > 1) of the default line")
    ap.add_argument("--place-reads", type=int, default=2_000_000)
    ap.add_argument("--mode-b-reads", type=int, default=4_000_000)
    ap.add_argument("--lanes", type=int, default=2, help="mode B: batches in flight per rank")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "c5":
            args.workload = "c3"  # same index, same reads: the reference holds the whole table in host memory
        reference_arm(args)
        return

    import krepp_b200
    env = Env(args.gpus)
    n = args.reads or DEFAULT_READS[args.workload]
    batch = min(args.batch or DEFAULT_BATCH[args.workload], n)
    t_wl = time.time()
    wl = Workload(args.workload, n, env.rank, env.world)
    t_wl = time.time() - t_wl
    cpu_sample = 0 if args.no_cpu_baseline else (args.cpu_sample or 200_000)

    if args.workload == "c5":  # mode B as the main line
        out = shard_arm(env, wl, n, batch, args.steps, args.warmup, not args.no_e2e, t_wl, cpu_sample, args.lanes)
    else:
        index = krepp_b200.Index(wl.index, env.local)
        wname = WORKLOADS[args.workload]
        place_name = wname.replace("configs[2]", "configs[3]").replace("krepp dist", "krepp place (per-read candidate placements: records of the jplace output)")
        out = hot_arm(env, index, wl, args.mode, n, batch, args.steps, args.warmup, args.e2e_batch, not args.no_e2e, cpu_sample,
                      place_name if args.mode == "place" else wname, t_wl)
        if args.mode == "dist" and args.workload == "c3" and not args.no_place:
            # configs[3] beside the headline: the same index and read pool through `krepp place`, shorter steps
            pn = min(args.place_reads, n)
            pl = hot_arm(env, index, wl, "place", pn, min(500_000, pn), args.steps, args.warmup, min(args.e2e_batch, 250_000), not args.no_e2e,
                         min(cpu_sample, 50_000), place_name, t_wl)
            if out is not None:
                out["place"] = pl
        index.close()
        if args.mode == "dist" and args.workload == "c3" and env.world > 1 and not args.no_mode_b:
            # configs[4] beside the headline: the same table split by bucket range over the ranks under a memory budget it exceeds
            bn = min(args.mode_b_reads, n)
            mb = shard_arm(env, wl, bn, min(1_000_000, bn), args.steps, args.warmup, not args.no_e2e, t_wl, min(cpu_sample, 50_000), args.lanes)
            if out is not None:
                out["mode_b"] = mb
    if env.rank == 0 and out is not None:
        print(json.dumps(out))
    env.close()


if __name__ == "__main__":
    main()
