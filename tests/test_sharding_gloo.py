"""The N>1 path on CPU: two gloo ranks shard the read stream the way bench.py / the CLI do across GPUs (contiguous
slices, replicated index, no data-path collective), and the only collectives -- max-over-ranks timing, counter sums,
rank-ordered output gathering -- behave.  The per-read compute stands in with the oracle here (CPU box, tests only)."""
import os
import sys

import pytest
import torch.multiprocessing as mp

from conftest import GOLDEN_DIR, ROOT

SMALL = os.path.join(GOLDEN_DIR, "small")


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import oracle_lib as O
    from krepp_b200 import dist as kd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with open(os.path.join(SMALL, "reads.fq"), "rb") as f:
            lines = f.read().split(b"\n")
        names = [lines[i][1:].decode() for i in range(0, len(lines) - 1, 4)]
        reads = [lines[i] for i in range(1, len(lines), 4)]
        b, e = kd.shard_range(len(reads), *kd.env_rank_world()[:2])
        ix = O.OracleIndex(os.path.join(SMALL, "index"))
        text, nrec = [], 0
        for nm, s in zip(names[b:e], reads[b:e]):
            o = ix.query(s)
            nrec += len(o["sel"])
            text += [f"{nm}\t{ix.name(x['leaf_se'])}\t{x['d']:.5f}\n" for x in o["sel"]] or [f"{nm}\tNA\tNaN\n"]
        tmax = kd.max_over_ranks(1.0 + rank)
        tot = kd.sum_over_ranks([e - b, nrec])
        blocks = kd.gather_text("".join(text))
        if rank == 0:
            q.put((tmax, tot, "".join(blocks)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_ranges_tile_the_input():
    from krepp_b200.dist import shard_range
    for n in (0, 1, 7, 236, 1_000_003):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


def test_two_rank_gloo_read_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, tot, text = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0            # max over ranks, not rank 0's own clock
    assert tot[0] == 236          # every read processed exactly once
    with open(os.path.join(SMALL, "ref_dist.tsv")) as f:
        ref = f.read().splitlines()
    assert sorted(text.splitlines()) == sorted(ref)   # concatenated per-rank output == the reference's output
    assert tot[1] == sum(1 for l in ref if not l.endswith("NaN"))
