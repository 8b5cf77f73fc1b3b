#!/usr/bin/env python
"""Host stage of `krepp index` (krepp_builder_set_union + krepp_builder_write: reference sets -> colour record -> files) at the
size of configuration 3, on the CPU alone: the union is reconstructed from a library tools/synth_index wrote (keys from inc-* /
the encoding column, one reference set per colour id in use, its leaves by expanding the colour), handed to the writer, and the
library written must equal the generator's: offsets and encodings byte for byte, the same number of colour ids, colours equal
by expansion on a sample.   usage: host_stage_scale.py <synth_index out dir>"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import krepp_b200  # noqa: E402
from libraries import colour_leaves, read_library, sampled_colour_check  # noqa: E402

src = sys.argv[1]
t0 = time.time()
lib = read_library(os.path.join(src, "index"))
cache = os.path.join(src, "union_cache.npz")
if os.path.exists(cache):  # (a second run times the writer alone)
    z = np.load(cache)
    keys, set_of, set_begin, set_leaves, rho = z["keys"], z["set_of"], z["set_begin"], z["set_leaves"], z["rho"]
    print(f"union of {src} from {cache}: {len(keys)} k-mers, {len(set_begin) - 1} reference sets", flush=True)
else:
    inc = np.frombuffer(lib["inc_bytes"], "<u8", lib["nrows"], 4)
    rows = np.repeat(np.arange(lib["nrows"], dtype=np.uint64), np.diff(inc, prepend=np.uint64(0)).astype(np.int64))
    keys = (rows << np.uint64(32)) | lib["enc"].astype(np.uint64)
    del rows
    used, set_of = np.unique(lib["se"], return_inverse=True)
    set_of = set_of.astype(np.uint32)
    exp = colour_leaves(lib)
    nnodes = lib["nnodes"]
    leaf_se = np.array([c for c in range(1, nnodes) if int(lib["pse"][c][0]) == 0 and int(lib["pse"][c][1]) == c], np.int64)
    rank_of = np.full(nnodes, -1, np.int64)
    rank_of[leaf_se] = np.arange(len(leaf_se))
    sizes = np.array([len(exp[int(c)]) for c in used], np.uint64)
    set_begin = np.zeros(len(used) + 1, np.uint64)
    set_begin[1:] = np.cumsum(sizes)
    set_leaves = np.empty(int(set_begin[-1]), np.uint32)
    for i, c in enumerate(used):
        a = int(set_begin[i])
        s = np.sort(rank_of[np.fromiter(exp[int(c)], np.int64)])
        set_leaves[a:a + len(s)] = s
    del exp
    rho = lib["rho"][leaf_se]
    np.savez(cache, keys=keys, set_of=set_of, set_begin=set_begin, set_leaves=set_leaves, rho=rho)
    print(f"reconstructed the union of {src}: {len(keys)} k-mers, {len(set_begin) - 1} reference sets, {int(set_begin[-1])} leaves in them ({time.time() - t0:.0f} s of Python)", flush=True)
names = lib["reflist"].decode().split()
g = krepp_b200.Index.geometry(lib["k"], lib["w"], lib["h"], lib["m"], lib["r"], bool(lib["frac"]), device=-1)
b = krepp_b200.LibraryBuilder(g, lib["tree"].decode(), names)
t1 = time.time()
b.set_union(keys, set_of, set_begin, set_leaves, rho)
t2 = time.time()
out = os.path.join(src, "host_index")
nk, nsub = b.write(out)
t3 = time.time()
b.close()
t4 = time.time()
print(f"krepp_builder_set_union (copies) {t2 - t1:.2f} s, krepp_builder_write {t3 - t2:.2f} s, destroy {t4 - t3:.2f} s on {os.cpu_count()} cores: {nk} k-mers, {nsub} colour ids")
mine = read_library(out)
print("offsets identical:", mine["inc_bytes"] == lib["inc_bytes"], " encoding column identical:", bool((mine["enc"] == lib["enc"]).all()), " rho identical:",
      mine["rho"].tobytes() == lib["rho"].tobytes(), " colour ids:", mine["nsubsets"], "generator", lib["nsubsets"])
n, bad = sampled_colour_check(mine, lib)
print(f"colours of {n} sampled k-mers expanded in both libraries: {bad} differ")
