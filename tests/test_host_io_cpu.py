"""CPU tests of the host I/O layer of the C-ABI library (krepp_reader_* / krepp_format_*): record framing against the
oracle's kseq restatement and (where built) the reference's own QSeq, text output against the golden files the
UNMODIFIED reference produced.  No GPU: the index handle is opened with KREPP_DEVICE_NONE and the result structs are
filled from the oracle's outputs."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

import conftest
from conftest import GOLDEN_DIR, needs_ref

SMALL = os.path.join(GOLDEN_DIR, "small")

TRICKY = [
    # (label, file text)
    ("fastq_plain", b"@r1 comment here\nACGTACGT\n+\nIIIIIIII\n@r2\nGGGGCCCC\n+r2\nJJJJJJJJ\n"),
    ("fastq_no_trailing_newline", b"@r1\nACGT\n+\nIIII\n@r2\nTTTT\n+\nIIII"),
    ("fastq_quality_starts_with_at", b"@r1\nACGTAC\n+\n@IIIII\n@r2\nGGTTAA\n+\n@@@@@@\n"),
    ("fastq_truncated_quality", b"@r1\nACGT\n+\nIIII\n@r2\nTTTTTTTT\n+\nIII"),
    ("fasta_multiline", b">c1 desc\nACGTAC\nGTACGT\n\nACG\n>c2\tdesc\nTTTT\n>c3\n>c4\nAC GT\n"),
    ("fasta_crlf_lower", b">c1\r\nacgtn\r\nACGT\r\n>c2\r\nGG\r\n"),
    ("fasta_nonprintable_dropped", b">c1\nAC\x01GT\xc8AC\x7fGT\n"),
    ("leading_garbage", b"garbage line\n\n>c1\nACGT\n"),
    ("mixed_fasta_fastq", b">c1\nACGT\n@r1\nGGGG\n+\nIIII\n>c2\nTT\n"),
    ("header_only_at_eof", b">c1\nACGT\n>c2"),
    ("empty_name", b">\nACGT\n> desc\nGG\n"),
    ("plus_inside_fasta_sequence", b">c1\nAC+GT\nIIII\n>c2\nAA\n"),
    ("empty_file", b""),
    ("empty_sequence_fastq", b"@r1\n\n+\n\n@r2\nAC\n+\nII\n"),
]


@pytest.fixture(scope="module")
def K():
    import krepp_b200
    krepp_b200.build_library()
    from krepp_b200 import capi
    return capi


def oracle_parse(text: bytes):
    import ctypes as C
    import oracle_lib as O
    L = O.lib()

    class Rec(C.Structure):
        _fields_ = [("name", C.c_char_p), ("seq", C.c_char_p), ("len", C.c_uint64)]
    L.ko_parse_reads.restype = C.c_int64
    L.ko_parse_reads.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(Rec))]
    recs = C.POINTER(Rec)()
    n = L.ko_parse_reads(text, len(text), C.byref(recs))
    return [recs[i].name.decode() for i in range(n)], [C.string_at(recs[i].seq, recs[i].len) for i in range(n)]


@pytest.mark.parametrize("label,text", TRICKY, ids=[t[0] for t in TRICKY])
def test_reader_framing_equals_oracle_kseq(K, tmp_path, label, text):
    p = tmp_path / "in.txt"
    p.write_bytes(text)
    names, reads = K.Reader(str(p)).read_all()
    onames, oreads = oracle_parse(text)
    assert names == onames and reads == oreads


def test_reader_known_answers(K, tmp_path):
    p = tmp_path / "in.fq"
    p.write_bytes(dict(TRICKY)["fastq_quality_starts_with_at"])
    assert K.Reader(str(p)).read_all() == (["r1", "r2"], [b"ACGTAC", b"GGTTAA"])
    p.write_bytes(dict(TRICKY)["fasta_multiline"])
    assert K.Reader(str(p)).read_all() == (["c1", "c2", "c3", "c4"], [b"ACGTACGTACGTACG", b"TTTT", b"", b"ACGT"])
    p.write_bytes(dict(TRICKY)["fastq_truncated_quality"])
    assert K.Reader(str(p)).read_all() == (["r1"], [b"ACGT"])
    p.write_bytes(dict(TRICKY)["fasta_nonprintable_dropped"])
    assert K.Reader(str(p)).read_all() == (["c1"], [b"ACGTACGT"])


# the reference itself crashes (std::string from a null kseq buffer) when the FIRST record of a file has no sequence
_REF_CASES = [t for t in TRICKY if t[0] not in ("empty_file", "empty_sequence_fastq")]


@needs_ref
@pytest.mark.parametrize("label,text", _REF_CASES, ids=[t[0] for t in _REF_CASES])
def test_reader_framing_equals_reference_qseq(K, tmp_path, label, text):
    """The reference's own QSeq/kseq (through oracle/_ref/ref_dump 'R' lines: name and length of every record)."""
    p = tmp_path / "in.txt"
    p.write_bytes(text)
    out = subprocess.run([os.path.join(conftest.REF_DIR, "ref_dump"), os.path.join(SMALL, "index"), str(p)], capture_output=True, text=True, check=True).stdout
    ref = [(t[2] if len(t) == 9 else "", int(t[-6])) for t in (l.split() for l in out.splitlines()) if t and t[0] == "R"]
    names, reads = K.Reader(str(p)).read_all()
    assert [(n, len(r)) for n, r in zip(names, reads)] == ref


def test_reader_gzip_and_batch_boundaries(K, tmp_path):
    """gzip input; batches cut by read count, base count and name bytes keep every record, in order; a record that does
    not fit opens the next batch; a record larger than the buffers is a capacity error."""
    with open(os.path.join(SMALL, "reads.fq"), "rb") as f:
        text = f.read()
    gz = tmp_path / "reads.fq.gz"
    with gzip.open(gz, "wb") as f:
        f.write(text)
    want = oracle_parse(text)
    assert K.Reader(str(gz)).read_all() == want
    assert K.Reader(str(gz)).read_all(max_reads=7) == want
    assert K.Reader(str(gz)).read_all(max_reads=1000, max_bases=6000) == want
    assert K.Reader(str(gz)).read_all(max_reads=50, max_name_bytes=40) == want
    r = K.Reader(str(gz))
    n1, r1, eof = r.next_batch(max_reads=1000, max_bases=6000)
    assert not eof and sum(map(len, r1)) <= 6000 < sum(map(len, r1)) + len(want[1][len(r1)])
    with pytest.raises(K.KreppError) as e:
        K.Reader(str(gz)).read_all(max_bases=100)
    assert e.value.code == 4


# ---------------------------------------------------------------------------------------------------- formatting

def results_from_oracle(K, outs, th=4):
    """krepp_results_t arrays in the layout the GPU path produces, filled from OracleIndex.query() dicts."""
    nrec = sum(len(o["minfo"]) for o in outs)
    npl = sum(len(o["place"]) for o in outs)
    reads = np.zeros(len(outs), K.READ_DTYPE)
    recs = np.zeros(nrec, K.RECORD_DTYPE)
    hist = np.zeros((nrec, th + 1), np.uint32)
    pls = np.zeros(npl, K.PLACEMENT_DTYPE)
    at = pat = 0
    for i, o in enumerate(outs):
        reads[i]["onmers"], reads[i]["wn"], reads[i]["hdist_filt"] = o["onmers"], o["wn"], o["hdist_filt"]
        reads[i]["rec_begin"], reads[i]["rec_count"], reads[i]["closest"] = at, len(o["minfo"]), -1
        sel = {(s["leaf_se"], s["strand"]): s for s in o["sel"]}
        for m in o["minfo"]:
            r = recs[at]
            r["read"], r["leaf_se"], r["strand"], r["match_count"], r["hdist_min"] = i, m["leaf_se"], m["strand"], m["match"], m["hdist_min"]
            r["rho"], r["d_llh"], r["v_llh"], r["chisq"], r["flags"] = m["rho"], m["d"], m["v"], np.nan, m["solved"]
            s = sel.get((m["leaf_se"], m["strand"]))
            if s is not None:
                r["flags"] |= 2 | (4 if s["is_closest"] else 0)
                r["chisq"] = s["chisq"]
                if s["is_closest"]:
                    reads[i]["closest"] = at
            hist[at] = m["hist"]
            at += 1
        reads[i]["place_begin"], reads[i]["place_count"], reads[i]["n_selected"] = pat, len(o["place"]), len(o["sel"])
        for q in o["place"]:
            pls[pat] = (i, q["se"], q["pendant"], q["distal"], -q["v"], q["lwr"], q["d"], q["chisq"])
            pat += 1
    return reads, recs, hist, pls


@pytest.fixture(scope="module")
def small(K):
    import krepp_b200
    import oracle_lib as O
    names, reads = K.Reader(os.path.join(SMALL, "reads.fq")).read_all()
    return dict(names=names, reads=reads, oracle=O.OracleIndex(os.path.join(SMALL, "index")), index=krepp_b200.Index(os.path.join(SMALL, "index"), device=-1))


def test_format_dist_equals_reference_tsv(K, small):
    import oracle_lib as O
    outs = [small["oracle"].query(s, O.default_params(no_filter=0)) for s in small["reads"]]
    arrs = results_from_oracle(K, outs)
    res = K.results_struct(*arrs)
    p = K.Params(4, 2.706, float("nan"), 2, 1, 1, 0, 0)
    got = K.format_dist(small["index"], p, res, small["names"])
    with open(os.path.join(SMALL, "ref_dist.tsv")) as f:
        ref = f.read().splitlines()
    assert sorted(got.splitlines()) == sorted(ref)
    hdr = K.format_header(small["index"], p, invocation="krepp dist -i x -q y")
    assert hdr.splitlines()[0].startswith("# software: krepp\tversion: ") and hdr.splitlines()[0].endswith("invocation :krepp dist -i x -q y")
    assert hdr.splitlines()[1] == "SEQ_ID\tREFERENCE_NAME\tDIST"
    # --no-multi: one line per read; --dist-max: NA rows for far reads; --filter keeps a subset
    nm = K.format_dist(small["index"], K.Params(4, 2.706, float("nan"), 2, 1, 0, 0, 0), res, small["names"]).splitlines()
    assert len(nm) == len(small["names"])
    dm = K.format_dist(small["index"], K.Params(4, 2.706, 0.05, 2, 1, 1, 0, 0), res, small["names"]).splitlines()
    assert all(l.endswith("NaN") or float(l.split("\t")[2]) < 0.05 + 1e-5 for l in dm) and len(dm) < len(ref)
    fl = K.format_dist(small["index"], K.Params(4, 2.706, float("nan"), 2, 0, 1, 0, 0), res, small["names"]).splitlines()
    assert set(fl) <= set(ref) and len(fl) < len(ref)
    # --summarize: weights of every read sum to 1 over its kept references
    w = np.zeros(small["index"].info.nnodes + 1)
    ps = K.Params(4, 2.706, float("nan"), 2, 1, 1, 1, 0)
    K.format_dist(small["index"], ps, res, small["names"], wcount=w)
    assert abs(w.sum() - sum(1 for o in outs if o["sel"])) < 1e-9
    foot = K.format_footer(small["index"], ps, wcount=w).splitlines()
    assert len(foot) == int((w > 0).sum()) and abs(sum(float(l.split("\t")[2]) for l in foot) - 1) < 1e-3
    # the 16-byte rows (krepp_brief_t, what `dist` front ends copy back) give the same text in every mode
    brief = K.brief_from_records(arrs[1], 2.706)
    rb = K.results_struct(arrs[0], None, None, brief=brief)
    for prm in (p, K.Params(4, 2.706, float("nan"), 2, 1, 0, 0, 0), K.Params(4, 2.706, 0.05, 2, 1, 1, 0, 0), K.Params(4, 2.706, float("nan"), 2, 0, 1, 0, 0),
                K.Params(4, 2.706, 0.08, 2, 0, 0, 0, 0)):
        assert K.format_dist(small["index"], prm, rb, small["names"]) == K.format_dist(small["index"], prm, res, small["names"])
    wb = np.zeros_like(w)
    K.format_dist(small["index"], ps, rb, small["names"], wcount=wb)
    assert np.array_equal(w, wb)
    # ... and so do the rows the device selects, orders and rounds itself (KREPP_OUT_DIST): 4 bytes per read + 4 per printed row
    for prm in (p, K.Params(4, 2.706, float("nan"), 2, 1, 0, 0, 0), K.Params(4, 2.706, 0.05, 2, 1, 1, 0, 0), K.Params(4, 2.706, float("nan"), 2, 0, 1, 0, 0),
                K.Params(4, 2.706, 0.08, 2, 0, 0, 0, 0), K.Params(4, 2.706, 0.03, 2, 0, 1, 0, 0)):
        db, dr = K.dist_rows_from_records(small["index"], prm, arrs[0], arrs[1])
        assert dr.dtype == np.uint32
        rc = K.results_struct(None, None, None, dist_begin=db, dist_rows=dr)
        assert K.format_dist(small["index"], prm, rc, small["names"]) == K.format_dist(small["index"], prm, res, small["names"])
    for prm in (ps, K.Params(4, 2.706, 0.05, 2, 1, 1, 1, 0)):
        db, dr = K.dist_rows_from_records(small["index"], prm, arrs[0], arrs[1])
        wf, wc = np.zeros_like(w), np.zeros_like(w)
        K.format_dist(small["index"], prm, res, small["names"], wcount=wf)
        K.format_dist(small["index"], prm, K.results_struct(None, None, None, dist_begin=db, dist_rows=dr), small["names"], wcount=wc)
        assert np.array_equal(wf, wc) and wf.sum() > 0


def test_format_place_equals_reference_jplace_on_untied_reads(K, small):
    import oracle_lib as O
    outs = [small["oracle"].query(s, O.default_params(want_place=1, no_filter=0)) for s in small["reads"]]
    arrs = results_from_oracle(K, outs)
    res = K.results_struct(*arrs)
    p = K.Params(4, 2.706, float("nan"), 2, 0, 1, 0, 1)
    inv = "krepp place -i x -q y"
    text = K.format_header(small["index"], p, invocation=inv) + K.format_place(small["index"], p, res, small["names"]) + \
        K.format_footer(small["index"], p, total_queries=len(outs), invocation=inv)
    jp = json.loads(text)
    with open(os.path.join(SMALL, "ref_place.jplace")) as f:
        ref = json.load(f)
    assert jp["tree"] == ref["tree"] and jp["fields"] == ref["fields"] and jp["version"] == ref["version"]
    assert jp["metadata"]["num_queries"] == ref["metadata"]["num_queries"]
    mine = {pl["n"][0]: pl["p"] for pl in jp["placements"]}
    theirs = {pl["n"][0]: pl["p"] for pl in ref["placements"]}
    checked = 0
    for name, o in zip(small["names"], outs):
        assert (name in mine) == bool(o["place"])
        if not o["sel"]:
            continue
        dmin = min(x["d"] for x in o["sel"])
        if sum(1 for x in o["sel"] if x["d"] == dmin) != 1:
            continue  # tied closest: the reference's own output varies between runs
        assert (name in mine) == (name in theirs), name
        if name in mine:
            assert sorted(map(tuple, mine[name])) == sorted(map(tuple, theirs[name])), name  # same 5-decimal text
            checked += 1
    assert checked > 100
    # raw text framing of one multi-candidate and one single-candidate read, as the reference writes it
    raw = K.format_place(small["index"], p, res, small["names"])
    assert '\t\t\t{"n" : ["' in raw and '"], "p" : [' in raw and "]\n\t\t\t}" in raw
    # --tabular rows and --no-multi
    tab = K.format_place(small["index"], p, res, small["names"], tabular=True).splitlines()
    assert len(tab) == len(arrs[3]) and all(len(l.split("\t")) == 5 for l in tab)
    pn = K.Params(4, 2.706, float("nan"), 2, 0, 0, 0, 1)
    one = K.format_place(small["index"], pn, res, small["names"], tabular=True).splitlines()
    assert len(one) == sum(1 for o in outs if o["place"])
    w = np.zeros(small["index"].info.nnodes + 1)
    K.format_place(small["index"], K.Params(4, 2.706, float("nan"), 2, 0, 1, 1, 1), res, small["names"], wcount=w)
    assert abs(w.sum() - len(one)) < 1e-9


def test_fixed5_equals_printf(K, small):
    """The fast fixed-5 path prints exactly what std::fixed << setprecision(5) prints, including near-ties."""
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.uniform(0, 0.5, 20000), rng.uniform(0, 1e-4, 2000), np.arange(0, 4000) * 1e-5 + 5e-6,
                           np.arange(0, 4000) * 1e-5 + 5e-6 + 1e-18, [0.0, 1.2625413546520176e-05, 0.5, 0.499995, 0.000005, 0.999995, 1e-10]])
    n = len(vals)
    reads = np.zeros(n, K.READ_DTYPE)
    recs = np.zeros(n, K.RECORD_DTYPE)
    hist = np.zeros((n, 5), np.uint32)
    reads["rec_begin"], reads["rec_count"], reads["closest"] = np.arange(n), 1, np.arange(n)
    recs["read"], recs["leaf_se"], recs["flags"], recs["d_llh"] = np.arange(n), 1, 7, vals
    res = K.results_struct(reads, recs, hist)
    names = ["q"] * n
    got = K.format_dist(small["index"], K.Params(4, 2.706, float("nan"), 2, 1, 1, 0, 0), res, names).splitlines()
    assert [l.split("\t")[2] for l in got] == ["%.5f" % v for v in vals]


def test_reader_fast_path_equals_state_machine(K, tmp_path, monkeypatch):
    """The four-line FASTQ fast path (memchr + range checks) must frame every input exactly as the byte-wise state machine
    (KREPP_READER_FAST=0): well-formed records mixed with everything that has to fall back -- wrapped sequence / quality
    lines, CRLF, comments, '@' / '+' / '>' inside quality strings, blank lines, FASTA records, short or missing quality,
    non-printable bytes -- and records cut by the 4 MB read buffer."""
    import random
    rng = random.Random(5)
    alpha = b"ACGTNacgtn"

    def seq(n):
        return bytes(rng.choice(alpha) for _ in range(n))

    def qual(n, nasty=False):
        pool = b"IIIIFFF#5:<" + (b"@+>" if nasty else b"")
        return bytes(rng.choice(pool) for _ in range(n))

    def record(i):
        n = rng.choice([0, 1, 7, 26, 27, 150, 150, 150, 151, 300])
        s, kind = seq(n), rng.randrange(12)
        name = b"r%d" % i
        if kind <= 4:
            return b"@" + name + b"\n" + s + b"\n+\n" + qual(n) + b"\n"
        if kind == 5:
            return b"@" + name + b" some comment\tx\n" + s + b"\n+" + name + b"\n" + qual(n, True) + b"\n"
        if kind == 6:
            return b"@" + name + b"\r\n" + s + b"\r\n+\r\n" + qual(n) + b"\r\n"
        if kind == 7:  # wrapped lines
            h = n // 2
            return b"@" + name + b"\n" + s[:h] + b"\n" + s[h:] + b"\n+\n" + qual(n)[:h] + b"\n" + qual(n)[h:] + b"\n"
        if kind == 8:
            return b">" + name + b" fasta\n" + s[:n // 3] + b"\n" + s[n // 3:] + b"\n"
        if kind == 9:
            return b"\n\n@" + name + b"\n" + s + b"\n+\n" + qual(max(n - 3, 0)) + b"\n"  # short quality: runs into what follows
        if kind == 10:
            return b"@" + name + b"\n" + s[:n // 2] + b" \x01" + s[n // 2:] + b"\n+\n" + qual(n) + b"\n"
        return b"@" + name + b"\n" + s + b"\n+\n" + qual(n) + b"x\n"  # one extra character after the quality string

    def parse(path, fast, **kw):
        monkeypatch.setenv("KREPP_READER_FAST", "1" if fast else "0")
        return K.Reader(str(path)).read_all(**kw)

    small = tmp_path / "mix.fq"
    small.write_bytes(b"".join(record(i) for i in range(3000)))
    a, b = parse(small, True), parse(small, False)
    assert a == b and len(a[0]) > 2500
    assert parse(small, True, max_reads=97, max_bases=4000) == b          # batch boundaries fall anywhere
    big = tmp_path / "big.fq"                                             # > 4 MB: records straddle the read buffer
    with open(big, "wb") as f:
        for i in range(30000):
            f.write(b"@q%d\n" % i + seq(150) + b"\n+\n" + qual(150) + b"\n" if i % 50 else record(i))
    a, b = parse(big, True), parse(big, False)
    assert a == b and len(a[0]) > 29000


def test_reader_chunk_parallel_equals_sequential(K, tmp_path, monkeypatch):
    """krepp_reader_set_threads: batches framed chunk-parallel from the mapped file hold exactly the records (and order) of the
    sequential reader -- clean four-line FASTQ with '@'-leading quality lines (the false record starts a chunk could land on),
    variable read lengths, a malformed stretch in the middle (wrapped records, FASTA) that only the state machine can walk, a
    truncated last record, and every batch limit (reads, bases, name bytes) cutting the batches."""
    import random
    rng = random.Random(11)

    def seq(n):
        return bytes(rng.choice(b"ACGTN") for _ in range(n))

    def qual(n):
        return bytes(rng.choice(b"@@@IIIIFFF#5:<+>") for _ in range(n))

    path = tmp_path / "par.fq"
    with open(path, "wb") as f:
        for i in range(120000):
            n = rng.choice([150, 150, 150, 151, 100, 27, 250])
            if 60000 <= i < 60040:   # a stretch the fast path cannot frame
                h = n // 2
                s, q = seq(n), qual(n)
                f.write(b"@w%d\n" % i + s[:h] + b"\n" + s[h:] + b"\n+\n" + q[:h] + b"\n" + q[h:] + b"\n" if i % 2 else b">fa%d\n" % i + seq(n) + b"\n")
            else:
                f.write(b"@r%d c\n" % i + seq(n) + b"\n+\n" + qual(n) + b"\n")
        f.write(b"@last\nACGTACGT\n+\nIIII")   # truncated quality: no record
    monkeypatch.setenv("KREPP_READER_FAST", "0")
    want = K.Reader(str(path)).read_all(max_reads=1 << 17, max_bases=1 << 25)
    monkeypatch.setenv("KREPP_READER_FAST", "1")
    assert len(want[0]) == 120000
    for threads in (2, 5, 8):
        for kw in (dict(max_reads=1 << 16, max_bases=1 << 24), dict(max_reads=40000, max_bases=3_000_000), dict(max_reads=1 << 16, max_bases=1 << 24, max_name_bytes=300_000),
                   dict(max_reads=3000, max_bases=1 << 22)):
            got = K.Reader(str(path), threads=threads).read_all(**kw)
            assert got == want, (threads, kw, len(got[0]))
