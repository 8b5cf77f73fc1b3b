"""GPU parity tests of mode B (include/krepp_b200.h "bucket-range shards"): the table split by LSH bucket
range over W ranks, lookups exchanged to the owning shard, hit entries exchanged back.  The results must be those of the
unsharded path bit for bit (integers AND doubles: the same records reach the same solve kernel) and agree with the
oracle.  W logical ranks share cuda:0 here (tensor hand-over instead of NCCL); the same phases over NCCL run in
test_two_gpu_nccl when the box has two GPUs."""
import os
import subprocess
import sys
import types

import numpy as np
import pytest

import conftest
from conftest import ROOT, TOY_DIR, needs_ref
from test_gpu_parity import fastq_reads

pytestmark = [pytest.mark.gpu]

SMALL = os.path.join(conftest.GOLDEN_DIR, "small")


def pack(reads):
    import torch
    from krepp_b200.capi import pack_reads
    bases, offs = pack_reads(reads)
    pad = np.zeros(len(bases) + 64, np.uint8)  # the kernels read 128-bit words
    pad[:len(bases)] = bases
    return torch.from_numpy(pad).cuda(), torch.from_numpy(offs.astype(np.int64)).cuda(), len(reads)


def assert_same_results(a: dict, b: dict, what=""):
    """Per read: same summary, same records in the same order with identical numbers, same histograms, same placements."""
    ra, rb = a["reads"], b["reads"]
    assert len(ra) == len(rb)
    for name in ("onmers", "wn", "hdist_filt", "rec_count", "place_count"):
        assert np.array_equal(ra[name], rb[name]), (what, "read summaries differ in", name)
    for i in range(len(ra)):
        ba, bb, n = int(ra["rec_begin"][i]), int(rb["rec_begin"][i]), int(ra["rec_count"][i])
        for name in a["records"].dtype.names:
            x, y = a["records"][name][ba:ba + n], b["records"][name][bb:bb + n]
            assert np.array_equal(x, y, equal_nan=(x.dtype.kind == "f")), (what, i, "records differ in", name, x, y)
        assert np.array_equal(a["hist"][ba:ba + n], b["hist"][bb:bb + n]), (what, i, "histograms differ")
        ca, cb = int(ra["closest"][i]), int(rb["closest"][i])
        assert (ca - ba if ca >= 0 else -1) == (cb - bb if cb >= 0 else -1), (what, i, "closest")
        pa, pb, pn = int(ra["place_begin"][i]), int(rb["place_begin"][i]), int(ra["place_count"][i])
        for name in a["placements"].dtype.names:
            x, y = a["placements"][name][pa:pa + pn], b["placements"][name][pb:pb + pn]
            assert np.array_equal(x, y, equal_nan=(x.dtype.kind == "f")), (what, i, "placements differ in", name)


def run_logical(index_dir, reads, world, **params):
    """Shards `reads` over `world` logical ranks on cuda:0 and runs them through mode B; returns per-rank (reads, results)."""
    import krepp_b200.dist as kd
    parts = [reads[slice(*kd.shard_range(len(reads), r, world))] for r in range(world)]
    cap_reads = max(max(len(p) for p in parts), 1)
    cap_bases = max(max(sum(len(s) for s in p) for p in parts), 1) + 64
    ranks = [kd.ShardRank(index_dir, 0, r, world, cap_reads, cap_bases, **params) for r in range(world)]
    job = kd.ShardedJob(ranks)
    res = job.run([pack(p) for p in parts])
    res = [{k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in r.items()} for r in res]
    info = dict(shards=[(int(r.index.shard.row0), int(r.index.shard.row1), int(r.index.shard.n_entries)) for r in ranks],
                nkmers=int(ranks[0].index.info.nkmers), exchanged=job.bytes_exchanged,
                alg=[r.slot.algorithmic_bytes() for r in ranks])
    for r in ranks:
        r.close()
    return parts, res, info


def unsharded(index_dir, reads, **params):
    import krepp_b200
    ix = krepp_b200.Index(index_dir, 0)
    os.environ["KREPP_PIPELINE"] = "sorted"
    try:
        b = krepp_b200.IBatch(ix, reads if len(reads) else [b""], **params)
    finally:
        del os.environ["KREPP_PIPELINE"]
    b.submit()
    r = b.wait()
    out = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in r.items()}
    alg = b.algorithmic_bytes()
    b.close()
    ix.close()
    return out, alg


@pytest.mark.parametrize("world", [2, 4])
def test_small_index_shards_match_unsharded_and_oracle(world):
    import oracle_lib as O
    from gpu_common import compare_read, gpu_stage_dicts
    idx = os.path.join(SMALL, "index")
    _, reads = fastq_reads(os.path.join(SMALL, "reads.fq"))
    parts, res, info = run_logical(idx, reads, world)
    assert sum(n for _, _, n in info["shards"]) == info["nkmers"]                      # the shards tile the table
    assert all(info["shards"][g][1] == info["shards"][g + 1][0] for g in range(world - 1))
    oracle = O.OracleIndex(idx)
    p = O.default_params(want_lookups=0)
    stats = dict(solves=0, bitexact_d=0, max_rel_d=0.0)
    alg = dict(bytes=0, lookups=0, entries=0)
    for part, r, a in zip(parts, res, info["alg"]):
        ref, _ = unsharded(idx, part)
        assert_same_results(r, ref, f"world {world}")
        g = gpu_stage_dicts(types.SimpleNamespace(n_reads=len(part)), r, None)
        for i, s in enumerate(part):
            o = oracle.query(s, p)
            o["minfo"] = [m for m in o["minfo"] if m["solved"]]  # the default output drops the pairs that fail the hdist_filt gate
            compare_read(i, g[i], o, False, False, stats)
        for k in alg:
            alg[k] += a[k]
    assert stats["solves"] > 500
    _, whole = unsharded(idx, reads)
    assert alg == whole  # SURVEY 8d bytes: summed over ranks they are those of the unsharded job (no shard is read twice)


def test_small_index_place_through_shards():
    idx = os.path.join(SMALL, "index")
    _, reads = fastq_reads(os.path.join(SMALL, "reads.fq"))
    kw = dict(place=True, no_filter=False)
    parts, res, _ = run_logical(idx, reads, 3, **kw)
    n = 0
    for part, r in zip(parts, res):
        ref, _ = unsharded(idx, part, **kw)
        assert_same_results(r, ref, "place")
        n += len(r["placements"])
    assert n == 408  # the golden jplace of tests/golden/small


def test_shard_edge_cases():
    """Empty batches on some ranks, reads shorter than k, N runs, more ranks than reads; buffers that must grow."""
    idx = os.path.join(SMALL, "index")
    _, reads = fastq_reads(os.path.join(SMALL, "reads.fq"))
    odd = [b"", b"ACGT", b"N" * 80, reads[0][:30] + b"NNNN" + reads[0][30:], reads[1].lower()]
    for rs, world in ((odd, 2), (reads[:3], 5), ([], 2)):
        parts, res, _ = run_logical(idx, rs, world)
        for part, r in zip(parts, res):
            if len(part):
                ref, _ = unsharded(idx, part)
                assert_same_results(r, ref, f"edge world {world}")
            else:
                assert len(r["reads"]) == 0 and len(r["records"]) == 0


def test_sharded_handle_refuses_the_unsharded_calls():
    import krepp_b200
    from krepp_b200.capi import KreppError
    ix = krepp_b200.Index(os.path.join(SMALL, "index"), 0, shard=1, nshards=2)
    b = krepp_b200.IBatch(ix, [b"ACGT" * 20])
    with pytest.raises(KreppError, match="shard"):
        b.submit()
    b.close()


@needs_ref
@pytest.mark.parametrize("world", [2, 3])
def test_toy_index_20k_reads_through_shards(world):
    import synth
    idx = os.path.join(TOY_DIR, "index_toy")
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    reads = [r.tobytes() for r in synth.sample_reads(seq, offs, 20000, seed=3)]
    parts, res, info = run_logical(idx, reads, world)
    ents = [n for _, _, n in info["shards"]]
    assert max(ents) - min(ents) < 0.01 * info["nkmers"]  # equal cmer bytes per shard
    for part, r in zip(parts, res):
        ref, _ = unsharded(idx, part)
        assert_same_results(r, ref, f"toy world {world}")


def test_two_gpu_nccl():
    """The same phases with a real NCCL all-to-all-v between two GPUs (skipped on one-GPU boxes)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "shard_nccl_worker.py")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "rank 0 ok" in p.stdout and "rank 1 ok" in p.stdout
