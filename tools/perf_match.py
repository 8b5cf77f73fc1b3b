#!/usr/bin/env python
"""Quick device-resident timing of the match kernel for both scan strategies (KREPP_SCAN=lane|staged).
usage: perf_match.py [n_reads] [strategies, comma separated]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import numpy as np
import torch
import synth
import krepp_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
groups = sys.argv[2].split(",") if len(sys.argv) > 2 else ["lane", "staged"]
toy = os.path.join(ROOT, "oracle", "_ref", "toy")
seq, offs = synth.load_packed(os.path.join(toy, "genomes.npz"))
reads = synth.sample_reads(seq, offs, n, seed=1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for g in groups:
    os.environ["KREPP_SCAN"] = g
    ix = krepp_b200.Index(os.path.join(toy, "index_toy"), 0)
    b = krepp_b200.IBatch(ix, reads)
    d_b = torch.from_numpy(b.bases).cuda()
    d_o = torch.from_numpy(b.offsets.astype(np.int64)).cuda()
    ms, tot = [], []
    for it in range(6):
        flush.zero_()
        torch.cuda.synchronize()
        b.submit_device(d_b.data_ptr(), d_o.data_ptr(), n, n * 150)
        r = b.wait()
        if it >= 2:
            ms.append(r["match_ms"]); tot.append(r["gpu_ms"])
    alg = b.algorithmic_bytes()
    print(f"scan={g:6s} match {np.mean(ms):7.3f} ms  total {np.mean(tot):7.3f} ms  -> {n / np.mean(ms) / 1e3:8.1f} M reads/s (match)  "
          f"{alg['bytes'] / np.mean(ms) / 1e6:7.1f} GB/s algorithmic  records {len(r['records'])}")
    b.close(); ix.close()
