"""CPU tests of mode B's host side: the bucket-range split the C ABI computes, and a two-rank gloo run of
the exchange plumbing (krepp_b200.dist.exchange_v) carrying the real payloads -- the oracle's lookups of each rank's
reads go to the rank that owns their bucket, that rank scans ITS slice of the table (numpy restatement of the
XOR/OR/popc filter, ref src/common.hpp:175, src/query.cpp:361-368), hit entries come back, and the home rank must hold
exactly the hits an unsharded scan finds.  No GPU and no compute call of the library is involved."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import GOLDEN_DIR, ROOT

SMALL = os.path.join(GOLDEN_DIR, "small", "index")
SFX = "-m4r1-frac"


def load_table(d):
    sfx = [f[8:] for f in os.listdir(d) if f.startswith("metadata-") and "." not in f][0]
    cm = np.fromfile(os.path.join(d, "cmer" + sfx), dtype=np.uint32, offset=8).reshape(-1, 2)
    inc = np.fromfile(os.path.join(d, "inc" + sfx), dtype=np.uint64, offset=4)
    return cm, inc


@pytest.mark.parametrize("nshards", [1, 2, 3, 8, 64])
def test_row_splits_tile_the_table_with_equal_bytes(nshards):
    import krepp_b200
    cm, inc = load_table(SMALL)
    at, prev_row = 0, 0
    sizes = []
    for g in range(nshards):
        ix = krepp_b200.Index(SMALL, device=-1, shard=g, nshards=nshards)
        s = ix.shard
        assert (s.shard, s.nshards) == (g, nshards)
        assert s.row0 == prev_row and s.first_entry == at            # contiguous, in order
        assert s.first_entry == (inc[s.row0 - 1] if s.row0 else 0) and s.first_entry + s.n_entries == (inc[s.row1 - 1] if s.row1 else 0)
        assert list(ix.row_splits) == list(krepp_b200.Index(SMALL, device=-1, shard=0, nshards=nshards).row_splits)  # same split on every rank
        at += s.n_entries
        prev_row = s.row1
        sizes.append(int(s.n_entries))
        ix.close()
    assert prev_row == len(inc) and at == len(cm) == int(inc[-1])
    biggest_bucket = int(np.diff(np.concatenate([[0], inc.astype(np.int64)])).max())
    assert max(sizes) - min(sizes) <= 2 * biggest_bucket + 1           # equal cmer bytes up to one bucket either side


def test_bad_shard_arguments():
    import krepp_b200
    from krepp_b200.capi import KreppError
    for shard, n in ((2, 2), (0, 0), (0, 257)):
        with pytest.raises(KreppError):
            krepp_b200.Index(SMALL, device=-1, shard=shard, nshards=n)


def _hd(enc, q):
    z = np.bitwise_xor(enc, q)
    z = (z | (z >> 16)) & 0xFFFF
    return np.array([bin(int(x)).count("1") for x in z], dtype=np.int64)


def _row(rix, m=4, r=1):
    return (rix // m) * (r + 1) + rix % m  # frac addressing, ref src/index.cpp:160-168


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import krepp_b200
    import oracle_lib as O
    from krepp_b200 import dist as kd
    from test_gpu_parity import fastq_reads
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        th = 4
        _, reads = fastq_reads(os.path.join(GOLDEN_DIR, "small", "reads.fq"))
        b, e = kd.shard_range(len(reads), rank, world)
        mine = reads[b:e]
        ix = krepp_b200.Index(SMALL, device=-1, shard=rank, nshards=world)  # metadata only: the split
        splits = [int(x) for x in ix.row_splits]
        cm, inc = load_table(SMALL)
        row0, row1, e0 = int(ix.shard.row0), int(ix.shard.row1), int(ix.shard.first_entry)
        my_cm = cm[e0:e0 + int(ix.shard.n_entries)]                          # this rank's slice of the table
        my_inc = inc[row0:row1].astype(np.int64) - e0
        oracle = O.OracleIndex(SMALL)
        p = O.default_params(want_lookups=1)
        tup = []  # {q, read, strand << 31 | lookup index, row}
        for i, s in enumerate(mine):
            for j, (strand, pos, rix, enc) in enumerate(oracle.query(s, p)["lookups"]):
                tup.append((enc, i, (strand << 31) | j, _row(rix)))
        tup = np.array(sorted(tup, key=lambda t: t[3]), dtype=np.int64).reshape(-1, 4)  # counting sort by row
        counts = [int(((tup[:, 3] >= splits[g]) & (tup[:, 3] < splits[g + 1])).sum()) for g in range(world)]
        recv, rc = kd.exchange_v(torch.from_numpy(tup), counts)
        recv = recv.numpy()
        assert len(recv) == sum(rc) and ((recv[:, 3] >= row0) & (recv[:, 3] < row1)).all()   # only rows this shard owns arrive
        hits, hcounts, at = [], [], 0
        for src in range(world):                                            # join, sender by sender
            n0 = len(hits)
            for enc, read, meta, row in recv[at:at + rc[src]]:
                lo = int(my_inc[row - row0 - 1]) if row > row0 else 0
                ents = my_cm[lo:int(my_inc[row - row0])]
                hd = _hd(ents[:, 0].astype(np.uint32), np.uint32(enc))
                for k in np.nonzero(hd <= th)[0]:
                    hits.append((read, meta, int(ents[k, 1]), int(hd[k])))
            at += rc[src]
            hcounts.append(len(hits) - n0)
        back, _ = kd.exchange_v(torch.tensor(hits, dtype=torch.int64).reshape(-1, 4), hcounts)
        got = sorted(map(tuple, back.numpy().tolist()))
        want = []                                                            # the unsharded scan of the same lookups
        for enc, read, meta, row in tup:
            lo = int(inc[row - 1]) if row else 0
            ents = cm[lo:int(inc[row])]
            hd = _hd(ents[:, 0], np.uint32(enc))
            want += [(int(read), int(meta), int(ents[k, 1]), int(hd[k])) for k in np.nonzero(hd <= th)[0]]
        assert got == sorted(want), (rank, len(got), len(want))
        tot = kd.sum_over_ranks([len(mine), len(tup), len(got)])
        if rank == 0:
            q.put(tot)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_bucket_range_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tot = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert tot[0] == 236 and tot[1] > 10_000 and tot[2] > 1_000


def _expected_checksums(d, row0=None, row1=None):
    """The sums krepp_index_host_checksums defines, from the index files alone (numpy + a plain restatement of the colour
    walk of ref src/query.cpp:369-387: a colour id is a tree leaf, or expands through crecord's pairs; null nodes drop out,
    a leaf reached twice counts once)."""
    import struct
    import krepp_b200
    sfx = [f[8:] for f in os.listdir(d) if f.startswith("metadata-") and "." not in f][0]
    cm = np.fromfile(os.path.join(d, "cmer" + sfx), dtype=np.uint64, offset=8)
    inc = np.fromfile(os.path.join(d, "inc" + sfx), dtype=np.uint64, offset=4)
    row0, row1 = (0, len(inc)) if row0 is None else (row0, row1)
    e0 = int(inc[row0 - 1]) if row0 else 0
    e1 = int(inc[row1 - 1]) if row1 else 0
    M = (1 << 64) - 1
    a = int(cm[e0:e1].sum(dtype=np.uint64))
    b = int((inc[row0:row1] - np.uint64(e0)).sum(dtype=np.uint64))
    raw = open(os.path.join(d, "crecord" + sfx), "rb").read()
    nn, ns = struct.unpack("<II", raw[:8])
    pse = np.frombuffer(raw[8:8 + 8 * ns], dtype="<u4").reshape(ns, 2)
    ix = krepp_b200.Index(d, device=-1)
    leaf = ix.tree()["is_leaf"]
    rank, r = {}, 0
    for se in range(1, nn):
        if leaf[se]:
            rank[se] = r
            r += 1
    memo = {}

    def leaves(c):
        if c in memo:
            return memo[c]
        if c == 0:
            out = frozenset()
        elif c < nn:
            out = frozenset([rank[c]]) if leaf[c] else leaves(int(pse[c][0])) | leaves(int(pse[c][1]))
        else:
            out = leaves(int(pse[c][0])) | leaves(int(pse[c][1]))
        memo[c] = out
        return out

    c_sum = d_sum = 0
    for c in range(ns):
        ls = leaves(c)
        c_sum = (c_sum + c * len(ls)) & M
        d_sum = (d_sum + (c + 1) * sum(l + 1 for l in ls)) & M
    return [a & M, b & M, c_sum, d_sum]


def test_host_image_checksums_whole_and_shards():
    """What the loader holds in memory -- the k-mer table (or a shard's slice of it), the rebased bucket ends and the flattened
    colour lists -- against the index files, without a device."""
    import sys
    import krepp_b200
    sys.setrecursionlimit(10000)
    from conftest import TOY_DIR, have_ref
    dirs = [SMALL] + ([os.path.join(TOY_DIR, "index_toy")] if have_ref() else [])   # 0.5 MB, and the 72 MB toy index
    for d in dirs:
        whole = krepp_b200.Index(d, device=-1)
        assert whole.host_checksums() == _expected_checksums(d), d
        for nshards in (2, 5):
            for g in range(nshards):
                ix = krepp_b200.Index(d, device=-1, shard=g, nshards=nshards)
                assert ix.host_checksums() == _expected_checksums(d, int(ix.shard.row0), int(ix.shard.row1)), (d, nshards, g)
