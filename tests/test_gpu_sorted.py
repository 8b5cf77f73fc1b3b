"""GPU parity tests of the bucket-sorted pipeline (krepp_b200/csrc/sorted.cu): the same stage-by-stage comparison with the
oracle as test_gpu_parity.py, with KREPP_PIPELINE=sorted, plus bit-identity with the fused kernel's output."""
import os

import numpy as np
import pytest

import conftest
from conftest import TOY_DIR, needs_ref
from test_gpu_parity import fastq_reads

pytestmark = [pytest.mark.gpu]


@pytest.fixture()
def sorted_pipeline(monkeypatch):
    monkeypatch.setenv("KREPP_PIPELINE", "sorted")


@pytest.fixture(scope="module")
def env():
    import krepp_b200
    import oracle_lib as O
    idx = os.path.join(TOY_DIR, "index_toy")
    return dict(dir=idx, oracle=O.OracleIndex(idx), gpu=krepp_b200.Index(idx, 0))


def test_sorted_golden_small_index(sorted_pipeline):
    import krepp_b200
    import oracle_lib as O
    from gpu_common import run_and_compare
    small = os.path.join(conftest.GOLDEN_DIR, "small")
    names, reads = fastq_reads(os.path.join(small, "reads.fq"))
    o, g = O.OracleIndex(os.path.join(small, "index")), krepp_b200.Index(os.path.join(small, "index"), 0)
    st = run_and_compare(small, reads, o, g)
    assert st["reads"] == 236 and st["solves"] > 500
    st = run_and_compare(small, reads, o, g, check_lookups=False, place=True, no_filter=False)
    assert st["placements"] == 408
    b = krepp_b200.IBatch(g, reads, names=names)
    with open(os.path.join(small, "ref_dist.tsv")) as f:
        assert sorted(b.estimate_distances().splitlines()) == sorted(f.read().splitlines())


@needs_ref
@pytest.mark.parametrize("lookup", ["binned", "two_pass"])
def test_sorted_toy_query_and_20k(env, sorted_pipeline, monkeypatch, lookup):
    """Both forms of the lookup sort: the two-pass counting sort (the default) and the two-level one (lookup_partition_kernel +
    bin_sort_kernel, KREPP_LOOKUP=binned)."""
    import synth
    from gpu_common import run_and_compare
    monkeypatch.setenv("KREPP_LOOKUP", lookup)
    names, reads = fastq_reads(os.path.join(TOY_DIR, "query_toy.fq"))
    st = run_and_compare(env["dir"], reads, env["oracle"], env["gpu"])
    assert st["reads"] == 100 and st["solves"] > 100
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    reads = [r.tobytes() for r in synth.sample_reads(seq, offs, 20000, seed=1)]
    st = run_and_compare(env["dir"], reads, env["oracle"], env["gpu"])
    assert st["max_rel_d"] < 1e-5


@needs_ref
@pytest.mark.parametrize("wide", ["0", "16"])
def test_sorted_edge_cases_and_long_reads(env, sorted_pipeline, monkeypatch, wide):
    """Empty / short / N / ragged reads, plus 20 kb reads whose leaf hits overflow the shared-memory sort buffer (the HBM
    scratch path); wide=16 pads the leaf field of the sort keys so that the 64-bit key path runs on this small index."""
    import synth
    from gpu_common import run_and_compare
    monkeypatch.setenv("KREPP_SORT_WIDE", wide)
    monkeypatch.setenv("KREPP_SEGMENT_WINDOWS", "0")  # whole reads: this test is about the one-warp-per-read paths (cut reads: test_gpu_segments.py)
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    rng = np.random.default_rng(5)
    base = [r.tobytes() for r in synth.sample_reads(seq, offs, 64, read_len=300, max_sub=0.05, seed=3)]
    reads = [b"", b"A", base[0][:26], base[0][:27], base[1][:28], b"N" * 150, base[2][:100].lower(),
             base[3][:60] + b"N" + base[3][61:150], b"ACGT" * 40, b"A" * 200, base[5][:149] + b"*", bytes([200]) + base[6][:150]]
    for ln in (127, 128, 129, 154, 155, 156, 255, 256, 257, 283, 300):
        reads.append(base[8 + (ln % 7)][:ln])
    reads += [r.tobytes() for r in synth.sample_reads(seq, offs, 4, read_len=20000, max_sub=0.02, seed=6)]
    reads += [r.tobytes() for r in synth.sample_reads(seq, offs, 2000, seed=8)]
    for _ in range(40):
        ln = int(rng.integers(1, 400))
        r = bytearray(base[int(rng.integers(0, 64))][:ln])
        for _ in range(int(rng.integers(0, 4))):
            r[int(rng.integers(0, len(r)))] = ord("N")
        reads.append(bytes(r))
    st = run_and_compare(env["dir"], reads, env["oracle"], env["gpu"])
    print(st)


@needs_ref
@pytest.mark.parametrize("th", [0, 2, 7, 9])
def test_sorted_other_thresholds(env, sorted_pipeline, th):
    import synth
    from gpu_common import run_and_compare
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    reads = [r.tobytes() for r in synth.sample_reads(seq, offs, 1500, seed=11 + th)]
    run_and_compare(env["dir"], reads, env["oracle"], env["gpu"], check_lookups=False, hdist_th=th)


@needs_ref
def test_sorted_place(env, sorted_pipeline):
    import synth
    from gpu_common import run_and_compare
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    reads = [r.tobytes() for r in synth.sample_reads(seq, offs, 4000, seed=77)]
    st = run_and_compare(env["dir"], reads, env["oracle"], env["gpu"], check_lookups=False, place=True, no_filter=False)
    assert st["placements"] > 500


@needs_ref
def test_sorted_equals_fused_bit_for_bit(env, monkeypatch):
    """Both pipelines must return the same per-read summaries and, read by read, the same records (order included),
    histograms and solved values; grown buffers (tiny initial capacity: one read per slot re-submitted with many) too."""
    import krepp_b200
    import synth
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    m = synth.sample_reads(seq, offs, 30000, seed=91)
    out = {}
    for pipe in ("fused", "sorted"):
        monkeypatch.setenv("KREPP_PIPELINE", pipe)
        b = krepp_b200.IBatch(env["gpu"], m)
        b.submit()
        r = b.wait()
        out[pipe] = {k: np.array(r[k], copy=True) for k in ("reads", "records", "hist")}
        out[pipe + "_alg"] = b.algorithmic_bytes()
        b.close()
    f, s = out["fused"], out["sorted"]
    assert out["fused_alg"] == out["sorted_alg"]
    for name in ("onmers", "wn", "hdist_filt", "rec_count"):
        assert np.array_equal(f["reads"][name], s["reads"][name]), name
    for i in range(len(f["reads"])):
        fb, sb, n = int(f["reads"]["rec_begin"][i]), int(s["reads"]["rec_begin"][i]), int(f["reads"]["rec_count"][i])
        for name in ("leaf_se", "strand", "match_count", "hdist_min", "flags", "d_llh", "v_llh"):
            assert np.array_equal(f["records"][name][fb:fb + n], s["records"][name][sb:sb + n], equal_nan=name in ("d_llh", "v_llh")), (i, name)
        assert np.array_equal(f["hist"][fb:fb + n], s["hist"][sb:sb + n]), (i, "hist")
        fc, sc = int(f["reads"]["closest"][i]), int(s["reads"]["closest"][i])
        assert (fc < 0 and sc < 0) or fc - fb == sc - sb, (i, "closest")


@needs_ref
def test_sorted_skewed_rows_fall_back_to_two_pass(env, sorted_pipeline, monkeypatch):
    """Reads that pile their lookups on a few rows (a thousand copies of poly-A, poly-AC and one genomic read among ordinary
    reads) overflow a coarse bin of the two-level lookup sort; the batch is redone with the exact two-pass sort and the slot stays
    on it: same records as the oracle either way, before and after."""
    import krepp_b200
    import synth
    from gpu_common import run_and_compare
    monkeypatch.setenv("KREPP_LOOKUP", "binned")
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    normal = [r.tobytes() for r in synth.sample_reads(seq, offs, 600, seed=17)]
    skew = [b"A" * 150] * 1000 + [b"AC" * 75] * 1000 + [normal[0]] * 1500
    run_and_compare(env["dir"], normal + skew, env["oracle"], env["gpu"])
    b = krepp_b200.IBatch(env["gpu"], normal + skew)
    b.submit(); b.wait()
    assert "lookup_kernel<scatter>" in [n for n, _ in b.stage_times()]       # the two-pass form ran ...
    b.bases, b.offsets = krepp_b200.capi.pack_reads(normal)
    b.n_reads = len(normal)
    b.submit(); r = b.wait()
    assert "lookup_kernel<scatter>" in [n for n, _ in b.stage_times()] and r["n_records"] > 500   # ... and the slot stays on it
    b.close()
    b = krepp_b200.IBatch(env["gpu"], normal)
    b.submit(); b.wait()
    assert "bin_sort_kernel" in [n for n, _ in b.stage_times()]              # a fresh slot (with KREPP_LOOKUP=binned) starts with the two-level form
    b.close()
