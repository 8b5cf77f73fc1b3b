"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same inputs.

Integer stages (k-mers/lookups, bucket ids, matched sets, Hamming histograms, filters, selections) must be bit-exact;
d_llh / v_llh / chisq within 1e-5 relative (north_star).  Needs oracle/_ref/toy (the reference-built toy index).
"""
import os

import numpy as np
import pytest

import conftest
from conftest import TOY_DIR, needs_ref

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def env():
    import krepp_b200
    import oracle_lib as O
    idx = os.path.join(TOY_DIR, "index_toy")
    return dict(dir=idx, oracle=O.OracleIndex(idx), gpu=krepp_b200.Index(idx, 0))


def fastq_reads(path):
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    return [lines[i][1:].split()[0].decode() for i in range(0, len(lines) - 1, 4)], [lines[i] for i in range(1, len(lines), 4)]


def test_golden_small_index_k21_h7():
    """The committed golden fixture (reference-built index with k=21 w=25 h=7, reads incl. edge cases): the CUDA path
    against the oracle on every stage, dist and place.  Needs nothing from oracle/_ref."""
    import krepp_b200
    import oracle_lib as O
    from gpu_common import run_and_compare
    small = os.path.join(conftest.GOLDEN_DIR, "small")
    names, reads = fastq_reads(os.path.join(small, "reads.fq"))
    o, g = O.OracleIndex(os.path.join(small, "index")), krepp_b200.Index(os.path.join(small, "index"), 0)
    st = run_and_compare(small, reads, o, g)
    assert st["reads"] == 236 and st["solves"] > 500
    st = run_and_compare(small, reads, o, g, check_lookups=False, place=True, no_filter=False)
    assert st["placements"] == 408  # the reference's own count for this fixture (ref_dump 'P' lines)
    b = krepp_b200.IBatch(g, reads, names=names)
    with open(os.path.join(small, "ref_dist.tsv")) as f:
        assert sorted(b.estimate_distances().splitlines()) == sorted(f.read().splitlines())
    for scan in ("lane", "staged"):   # every scan strategy gives the same integers
        os.environ["KREPP_SCAN"] = scan
        try:
            g2 = krepp_b200.Index(os.path.join(small, "index"), 0)
            run_and_compare(small, reads, o, g2, check_lookups=False)
            run_and_compare(small, reads, o, g2, check_lookups=False, place=True, no_filter=False)
            g2.close()
        finally:
            del os.environ["KREPP_SCAN"]


@needs_ref
def test_toy_query_all_stages(env):
    from gpu_common import run_and_compare
    names, reads = fastq_reads(os.path.join(TOY_DIR, "query_toy.fq"))
    st = run_and_compare(env["dir"], reads, env["oracle"], env["gpu"])
    assert st["reads"] == 100 and st["solves"] > 100
    print(st)


@needs_ref
def test_synthetic_20k_all_stages(env):
    import synth
    from gpu_common import run_and_compare
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    reads = [r.tobytes() for r in synth.sample_reads(seq, offs, 20000, seed=1)]
    st = run_and_compare(env["dir"], reads, env["oracle"], env["gpu"])
    print(st)
    assert st["max_rel_d"] < 1e-5          # north_star tolerance for distances
    assert st["bitexact_d"] >= 0.9 * st["solves"]  # device pow/log differ from glibc by an ulp now and then


@needs_ref
def test_edge_cases(env):
    """Empty / shorter-than-k / exactly-k reads, N runs, lower case, non-ACGT bytes, long and ragged reads."""
    import synth
    from gpu_common import run_and_compare
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    rng = np.random.default_rng(5)
    base = [r.tobytes() for r in synth.sample_reads(seq, offs, 64, read_len=300, max_sub=0.05, seed=3)]
    reads = [b"", b"A", base[0][:26], base[0][:27], base[1][:28], b"N" * 150, base[2][:100].lower(),
             base[3][:60] + b"N" + base[3][61:150], base[4][:30] + b"NNNNN" + base[4][35:200], b"ACGT" * 40, b"A" * 200,
             base[5][:149] + b"*", bytes([200]) + base[6][:150], base[7][:150].replace(b"A", b"R")]
    for ln in (127, 128, 129, 154, 155, 156, 255, 256, 257, 283, 300):
        reads.append(base[8 + (ln % 7)][:ln])
    long_src = synth.sample_reads(seq, offs, 4, read_len=5000, max_sub=0.02, seed=9)
    reads += [long_src[0].tobytes(), long_src[1].tobytes()[:1000], long_src[2].tobytes()[:3333]]
    for _ in range(40):
        ln = int(rng.integers(1, 400))
        r = bytearray(base[int(rng.integers(0, 64))][:ln])
        for _ in range(int(rng.integers(0, 4))):
            r[int(rng.integers(0, len(r)))] = ord("N")
        reads.append(bytes(r))
    st = run_and_compare(env["dir"], reads, env["oracle"], env["gpu"])
    print(st)


@needs_ref
@pytest.mark.parametrize("th", [0, 2, 7])
def test_other_thresholds(env, th):
    import synth
    from gpu_common import run_and_compare
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    reads = [r.tobytes() for r in synth.sample_reads(seq, offs, 1500, seed=11 + th)]
    run_and_compare(env["dir"], reads, env["oracle"], env["gpu"], check_lookups=False, hdist_th=th)


@needs_ref
def test_dist_filter_chisq(env):
    import synth
    from gpu_common import run_and_compare
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    reads = [r.tobytes() for r in synth.sample_reads(seq, offs, 3000, seed=21)]
    run_and_compare(env["dir"], reads, env["oracle"], env["gpu"], check_lookups=False, no_filter=False)


@needs_ref
@pytest.mark.parametrize("kw", [dict(no_filter=False), dict(no_filter=True), dict(no_filter=False, tau=3, chisq=3.841)])
def test_place_all_stages(env, kw):
    """krepp place: candidate edges identical to the oracle's (same deterministic tie rule), likelihoods / LWR / pendant
    lengths within 1e-5."""
    import synth
    from gpu_common import run_and_compare
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    names, reads = fastq_reads(os.path.join(TOY_DIR, "query_toy.fq"))
    reads = reads + [r.tobytes() for r in synth.sample_reads(seq, offs, 6000, seed=77)]
    st = run_and_compare(env["dir"], reads, env["oracle"], env["gpu"], check_lookups=False, place=True, **kw)
    print(st)
    assert st["placements"] > 1000


@needs_ref
def test_tsv_matches_reference_cli(env):
    """`krepp dist` body of the reference binary itself vs the GPU path's TSV, compared as sorted line sets."""
    import subprocess
    import krepp_b200
    names, reads = fastq_reads(os.path.join(TOY_DIR, "query_toy.fq"))
    ref = subprocess.run([os.path.join(conftest.REF_DIR, "krepp"), "dist", "-i", env["dir"], "-q", os.path.join(TOY_DIR, "query_toy.fq")],
                         capture_output=True, text=True, check=True).stdout.splitlines()[2:]
    b = krepp_b200.IBatch(env["gpu"], reads, names=names)
    got = b.estimate_distances().splitlines()
    assert sorted(got) == sorted(ref)


@needs_ref
def test_device_resident_input_and_idempotence(env):
    """submit_device on HBM-resident reads gives the same records as the host path; resubmitting a slot is idempotent."""
    import torch
    import krepp_b200
    import synth
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    m = synth.sample_reads(seq, offs, 5000, seed=33)
    b = krepp_b200.IBatch(env["gpu"], m)
    b.submit(); r1 = b.wait()
    key1 = (r1["reads"].copy(), np.sort(r1["records"].copy(), order=["read", "strand", "leaf_se"]))
    d_b = torch.from_numpy(b.bases.copy()).cuda()
    d_o = torch.from_numpy(b.offsets.astype(np.int64)).cuda()
    for _ in range(2):
        b.submit_device(d_b.data_ptr(), d_o.data_ptr(), b.n_reads, int(b.offsets[-1]))
        r2 = b.wait()
        rec2 = np.sort(r2["records"].copy(), order=["read", "strand", "leaf_se"])
        for f in ("leaf_se", "strand", "match_count", "hdist_min", "flags"):
            assert np.array_equal(key1[1][f], rec2[f]), f
        assert np.array_equal(key1[1]["d_llh"], rec2["d_llh"])
        for f in ("onmers", "wn", "hdist_filt", "rec_count"):
            assert np.array_equal(key1[0][f], r2["reads"][f]), f


# ---------------------------------------------------------------------------------------------------- the CLI (C++ host)

def _cli(*args):
    import subprocess
    exe = os.path.join(conftest.ROOT, "krepp_b200", "_build", "krepp_b200")
    return subprocess.run([exe, *args], capture_output=True, text=True)


def test_cli_dist_golden_small(tmp_path):
    """`krepp_b200 dist` (C++ host over the C ABI: reader -> GPU -> formatter) against the reference's own TSV for the
    committed fixture; tiny batches, several formatter threads and gzip input must not change a byte."""
    import gzip
    small = os.path.join(conftest.GOLDEN_DIR, "small")
    with open(os.path.join(small, "ref_dist.tsv")) as f:
        ref = sorted(f.read().splitlines())
    r = _cli("dist", "-i", os.path.join(small, "index"), "-q", os.path.join(small, "reads.fq"))
    assert r.returncode == 0, r.stderr
    body = r.stdout.splitlines()
    assert body[0].startswith("# software: krepp") and body[1] == "SEQ_ID\tREFERENCE_NAME\tDIST"
    assert sorted(body[2:]) == ref
    assert "Total number of sequences queried: 236" in r.stderr
    gz = tmp_path / "reads.fq.gz"
    with open(os.path.join(small, "reads.fq"), "rb") as f, gzip.open(gz, "wb") as g:
        g.write(f.read())
    out = tmp_path / "out.tsv"
    r2 = _cli("--num-threads", "3", "dist", "-i", os.path.join(small, "index"), "-q", str(gz), "-o", str(out), "--batch-reads", "17", "--slots", "2")
    assert r2.returncode == 0, r2.stderr
    assert out.read_text().splitlines()[2:] == body[2:]  # same rows in the same (input) order


def test_cli_shard_index_same_bytes(tmp_path):
    """--shard-index (SURVEY.md 8e mode B inside one process: every --devices entry holds one bucket-range shard, runs of lookups
    and hits move between them by peer copies) must not change a byte of dist / place output.  Three shards on cuda:0 here; on a
    multi-GPU box also one shard per GPU."""
    import torch
    small = os.path.join(conftest.GOLDEN_DIR, "small")
    base = ("-i", os.path.join(small, "index"), "-q", os.path.join(small, "reads.fq"))
    layouts = [("--devices", "0,0,0")] + ([("--num-gpus", "2")] if torch.cuda.device_count() >= 2 else [])
    for sub in (("dist",), ("dist", "--filter", "--batch-reads", "29"), ("place", "--batch-reads", "40"), ("place", "--tabular")):
        one = _cli("--num-threads", "2", *sub, *base)
        assert one.returncode == 0, one.stderr
        for lay in layouts:
            sh = _cli("--num-threads", "2", *sub, *base, "--shard-index", *lay)
            assert sh.returncode == 0, sh.stderr
            if sub[0] == "place" and "--tabular" not in sub:
                import json
                a, b = json.loads(one.stdout), json.loads(sh.stdout)
                assert a["placements"] == b["placements"] and a["tree"] == b["tree"] and len(b["placements"]) > 100
            else:
                assert one.stdout.splitlines()[1:] == sh.stdout.splitlines()[1:] and len(sh.stdout.splitlines()) > 200, (sub, lay)
            assert "Total number of sequences queried: 236" in sh.stderr
    if conftest.have_ref():  # and at a size where a missing synchronisation between the phases shows: 20,000 reads on the toy index
        import synth
        seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
        fq = str(tmp_path / "toy20k.fq")
        synth.write_fastq(fq, synth.sample_reads(seq, offs, 20000, seed=8))
        base = ("--num-threads", "4", "dist", "-i", os.path.join(TOY_DIR, "index_toy"), "-q", fq, "--batch-reads", "3000")
        one = _cli(*base)
        assert one.returncode == 0, one.stderr
        for lay in [("--devices", "0,0")] + layouts:
            sh = _cli(*base, "--shard-index", *lay)
            assert sh.returncode == 0, sh.stderr
            assert one.stdout.splitlines()[1:] == sh.stdout.splitlines()[1:] and len(sh.stdout.splitlines()) > 20000, lay


def test_cli_two_gpus_same_bytes(tmp_path):
    """--num-gpus 2 (index replicated, batches dealt round-robin to the GPUs: SURVEY.md 8e mode A) must not change a byte of
    the output, which stays in input order.  Skipped on one-GPU boxes."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    small = os.path.join(conftest.GOLDEN_DIR, "small")
    args = ("--num-threads", "2", "dist", "-i", os.path.join(small, "index"), "-q", os.path.join(small, "reads.fq"), "--batch-reads", "23")
    one, two = _cli(*args), _cli(*args, "--num-gpus", "2")
    assert one.returncode == 0 and two.returncode == 0, two.stderr
    assert one.stdout.splitlines()[1:] == two.stdout.splitlines()[1:] and len(two.stdout.splitlines()) > 200
    pl = ("place", "-i", os.path.join(small, "index"), "-q", os.path.join(small, "reads.fq"), "--batch-reads", "31", "--tabular")
    one, two = _cli(*pl), _cli(*pl, "--devices", "1,0")
    assert one.returncode == 0 and two.returncode == 0, two.stderr
    assert one.stdout.splitlines()[1:] == two.stdout.splitlines()[1:]


def test_cli_place_golden_small(tmp_path):
    import json
    import oracle_lib as O
    small = os.path.join(conftest.GOLDEN_DIR, "small")
    r = _cli("place", "-i", os.path.join(small, "index"), "-q", os.path.join(small, "reads.fq"), "--batch-reads", "50")
    assert r.returncode == 0, r.stderr
    jp = json.loads(r.stdout)
    with open(os.path.join(small, "ref_place.jplace")) as f:
        ref = json.load(f)
    assert jp["tree"] == ref["tree"] and jp["fields"] == ref["fields"] and jp["metadata"]["num_queries"] == "236"
    o = O.OracleIndex(os.path.join(small, "index"))
    names, reads = fastq_reads(os.path.join(small, "reads.fq"))
    p = O.default_params(want_place=1, no_filter=0)
    mine = {pl["n"][0]: pl["p"] for pl in jp["placements"]}
    nrows = 0
    for name, s in zip(names, reads):
        want = o.query(s, p)["place"]
        assert (name in mine) == bool(want), name
        if want:
            got = mine[name]
            assert [row[0] for row in got] == [q["edge"] for q in want], name
            for row, q in zip(got, want):
                for a, b in zip(row[1:], (q["pendant"], q["distal"], -q["v"], q["lwr"], q["d"])):
                    assert abs(a - b) <= 1.001e-5 + 1e-5 * abs(b), (name, row, q)
            nrows += len(got)
    assert nrows == 408
    t = _cli("place", "-i", os.path.join(small, "index"), "-q", os.path.join(small, "reads.fq"), "--tabular")
    assert t.returncode == 0 and len(t.stdout.splitlines()) == 408 + 3
    bad = _cli("place", "-i", os.path.join(small, "index"), "-q", os.path.join(small, "reads.fq"), "--tau", "5")
    assert bad.returncode != 0 and "Invalid configuration" in bad.stderr


@needs_ref
def test_cli_dist_equals_reference_cli_on_toy(env):
    import subprocess
    q = os.path.join(TOY_DIR, "query_toy.fq")
    ref = subprocess.run([os.path.join(conftest.REF_DIR, "krepp"), "dist", "-i", env["dir"], "-q", q], capture_output=True, text=True, check=True).stdout.splitlines()[2:]
    r = _cli("dist", "-i", env["dir"], "-q", q)
    assert r.returncode == 0, r.stderr
    assert sorted(r.stdout.splitlines()[2:]) == sorted(ref)


@needs_ref
@pytest.mark.parametrize("scan", ["lane", "staged"])
def test_scan_strategies_on_toy(scan):
    """Both phase-B strategies of the match kernel (lane-per-bucket, bulk-copy ring) on the toy index: all integer stages
    bit-exact, incl. long reads that span many tiles and ring wrap-arounds."""
    import krepp_b200
    import oracle_lib as O
    import synth
    from gpu_common import run_and_compare
    idx = os.path.join(TOY_DIR, "index_toy")
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    reads = [r.tobytes() for r in synth.sample_reads(seq, offs, 4000, seed=5)]
    reads += [r.tobytes() for r in synth.sample_reads(seq, offs, 6, read_len=20000, max_sub=0.03, seed=6)]
    os.environ["KREPP_SCAN"] = scan
    try:
        g = krepp_b200.Index(idx, 0)
        st = run_and_compare(idx, reads, O.OracleIndex(idx), g)
        assert st["max_rel_d"] < 1e-5
        g.close()
    finally:
        del os.environ["KREPP_SCAN"]


def test_brief_rows_equal_full_rows():
    """KREPP_OUT_BRIEF: the 16-byte rows the `dist` front end copies back carry the same reference, strand, flags, filter
    decision and distance (bit for bit) as the full records, and format to the same TSV."""
    import krepp_b200
    from krepp_b200 import capi
    small = os.path.join(conftest.GOLDEN_DIR, "small")
    names, reads = fastq_reads(os.path.join(small, "reads.fq"))
    g = krepp_b200.Index(os.path.join(small, "index"), 0)
    for kw in (dict(), dict(no_filter=False)):
        b = krepp_b200.IBatch(g, reads, names=names, **kw)
        b.set_output(brief=True)
        b.submit()
        r = {k: np.array(v, copy=True) for k, v in b.wait().items() if isinstance(v, np.ndarray)}  # the views die with the next submit
        want = capi.brief_from_records(r["records"], 2.706)
        assert len(r["brief"]) == len(r["records"]) > 500
        for f in ("read", "ref", "d_llh"):
            assert np.array_equal(r["brief"][f], want[f]), f
        b.set_output(records=False, hist=False, placements=False, brief=True)   # what the CLI asks for
        b.submit()
        r2 = {k: np.array(v, copy=True) for k, v in b.wait().items() if isinstance(v, np.ndarray)}
        assert len(r2["records"]) == 0 and all(np.array_equal(r2["reads"][f], r["reads"][f]) for f in ("onmers", "wn", "hdist_filt", "rec_count"))
        # (reads land in the row arrays in completion order, which differs from run to run: compare per read)
        for i in range(len(reads)):
            a0, n0, a1 = int(r["reads"]["rec_begin"][i]), int(r["reads"]["rec_count"][i]), int(r2["reads"]["rec_begin"][i])
            assert np.array_equal(r["brief"][a0:a0 + n0], r2["brief"][a1:a1 + n0]), i
        full = capi.results_struct(r["reads"], r["records"], r["hist"])
        brief = capi.results_struct(r2["reads"], None, None, brief=r2["brief"])
        assert capi.format_dist(g, b.params, full, names) == capi.format_dist(g, b.params, brief, names)
        b.close()


def test_dist_rows_selected_on_the_device():
    """KREPP_OUT_DIST: the rows `krepp dist` prints, selected / ordered / rounded by the device, equal the selection made on the
    host from the full records (a restatement of report_distances, ref src/query.cpp:158-196) in every mode, format to the same
    TSV, and are all that leaves the device when asked for alone; krepp_batch_set_output refuses a pending slot."""
    import math
    import krepp_b200
    from krepp_b200 import capi
    small = os.path.join(conftest.GOLDEN_DIR, "small")
    names, reads = fastq_reads(os.path.join(small, "reads.fq"))
    g = krepp_b200.Index(os.path.join(small, "index"), 0)
    modes = [dict(), dict(no_filter=False), dict(multi=False), dict(dist_max=0.05), dict(no_filter=False, dist_max=0.03), dict(multi=False, dist_max=0.08),
             dict(summarize=True), dict(summarize=True, dist_max=0.05)]
    for kw in modes:
        b = krepp_b200.IBatch(g, reads, names=names, **kw)
        b.set_output(dist=True)
        b.submit()
        with pytest.raises(capi.KreppError):
            b.set_output(dist=False)          # a batch is pending
        r = {k: np.array(v, copy=True) for k, v in b.wait().items() if isinstance(v, np.ndarray)}
        want_begin, want_rows = capi.dist_rows_from_records(g, b.params, r["reads"], r["records"])
        assert np.array_equal(r["dist_begin"], want_begin), kw
        assert np.array_equal(r["dist_rows"], want_rows) and r["dist_rows"].dtype == np.uint32, kw
        if not kw.get("summarize"):
            assert len(want_rows) > (100 if kw.get("dist_max") else 200) or kw.get("multi") is False
            full = capi.results_struct(r["reads"], r["records"], r["hist"])
            compact = capi.results_struct(None, None, None, dist_begin=r["dist_begin"], dist_rows=r["dist_rows"])
            assert capi.format_dist(g, b.params, full, names) == capi.format_dist(g, b.params, compact, names)
        b.set_output(records=False, hist=False, placements=False, summaries=False, dist=True)   # what the dist command line asks for
        b.submit()
        r2 = b.wait()
        assert len(r2["records"]) == 0 and len(r2["reads"]) == 0 and len(r2["brief"]) == 0
        assert np.array_equal(r2["dist_begin"], want_begin) and np.array_equal(r2["dist_rows"], want_rows), kw
        sums = b.wait_device()["reads"]       # the summaries are still to be had
        assert np.array_equal(sums["onmers"], r["reads"]["onmers"]) and np.array_equal(sums["rec_count"], r["reads"]["rec_count"])
        b.close()
    # empty batch
    b = krepp_b200.IBatch(g, [b""], capacity=(4, 64))
    b.set_output(records=False, hist=False, placements=False, summaries=False, dist=True)
    b.submit()
    r = b.wait()
    assert list(r["dist_begin"]) == [0x80000000, 0] and len(r["dist_rows"]) == 0
    b.close()
