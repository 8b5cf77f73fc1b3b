"""Benchmark workloads (measurement infrastructure; nothing here is on the product path).

  toy : BASELINE.json configs[1] -- the reference-built toy index (oracle/_ref/toy, made by `make -C oracle toy` where the
        reference sources exist) and synthetic 150 bp reads sampled from the toy genomes (tools/synth.py).
  c3  : BASELINE.json configs[2] -- 1,000 synthetic genomes x 3 Mbp on a random binary tree, krepp-format index, 150 bp
        reads with 0-15 % substitutions.  Nothing persists on a GPU box and a 1.6 GB index cannot be shipped, so it is
        generated where the benchmark runs by tools/_build/synth_index (about half a minute on 16 cores) into a cache
        directory under /tmp that every rank and both arms of the same box share (file lock; rank 0 of a torchrun job
        and the reference arm find it already built).
"""
from __future__ import annotations

import fcntl
import json
import os
import subprocess
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOY = os.path.join(ROOT, "oracle", "_ref", "toy")
SYNTH = os.path.join(ROOT, "tools", "_build", "synth_index")
CACHE = os.environ.get("KREPP_WORKLOAD_CACHE", "/tmp/krepp_b200_workloads")
READ_LEN = 150


def ensure_c3(total_reads: int, genomes: int = 1000, length: int = 3_000_000, seed: int = 7, fastq_reads: int = 200_000) -> tuple[str, dict]:
    """Returns (directory, workload.json) of the config-3 workload with `total_reads` reads, building it if needed."""
    if not os.path.exists(SYNTH):
        raise RuntimeError(f"{SYNTH} is missing: run __graft_entry__.build()")
    d = os.path.join(CACHE, f"c3_g{genomes}_l{length}_s{seed}_r{total_reads}")
    os.makedirs(CACHE, exist_ok=True)
    with open(d + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not os.path.exists(os.path.join(d, "workload.json")):
                t0 = time.time()
                subprocess.run([SYNTH, "--out", d, "--genomes", str(genomes), "--length", str(length), "--seed", str(seed), "--reads", str(total_reads),
                                "--fastq-reads", str(min(fastq_reads, total_reads))], check=True, stdout=subprocess.DEVNULL)
                with open(os.path.join(d, "built_s.txt"), "w") as f:
                    f.write(f"{time.time() - t0:.1f}\n")
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    with open(os.path.join(d, "workload.json")) as f:
        return d, json.load(f)


def c3_reads(d: str, first: int, n: int) -> np.ndarray:
    """Reads [first, first + n) of the pool as an (n, 150) uint8 matrix.  The generator scatters reads through a
    multiplicative permutation, so any contiguous slice is a uniform sample over genomes, strands and error rates."""
    m = np.memmap(os.path.join(d, "reads.u8"), dtype=np.uint8, mode="r")
    return np.ascontiguousarray(m[first * READ_LEN:(first + n) * READ_LEN]).reshape(n, READ_LEN)


def toy_reads(n: int, seed: int) -> np.ndarray:
    import synth
    seq, offs = synth.load_packed(os.path.join(TOY, "genomes.npz"))
    return synth.sample_reads(seq, offs, n, read_len=READ_LEN, seed=seed)
