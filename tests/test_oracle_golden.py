"""CPU tests: the plain-C oracle against the committed golden fixtures (tests/golden/small), which were produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  This is what pins the oracle on a box without /root/reference."""
import gzip
import json
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

SMALL = os.path.join(GOLDEN_DIR, "small")


def read_fastq(path):
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    return [lines[i][1:].split()[0].decode() for i in range(0, len(lines) - 1, 4)], [lines[i] for i in range(1, len(lines), 4)]


@pytest.fixture(scope="module")
def golden():
    import oracle_lib as O
    with gzip.open(os.path.join(SMALL, "ref_dump.txt.gz"), "rt") as f:
        dump = O.parse_ref_dump(f.read())
    names, reads = read_fastq(os.path.join(SMALL, "reads.fq"))
    return dict(dump=dump, names=names, reads=reads, oracle=O.OracleIndex(os.path.join(SMALL, "index")))


def test_index_geometry(golden):
    import oracle_lib as O
    info, ix = golden["dump"]["info"], golden["oracle"]
    assert (ix.k, ix.hh, ix.m, ix.nnodes) == (info["k"], info["h"], info["m"], info["nnodes"]) == (21, 7, 4, 15)
    assert O.lib().ko_index_mask_hash_bp(ix.h) == info["mask_hash_bp"]
    assert O.lib().ko_index_mask_drop_lr(ix.h) == info["mask_drop_lr"]


def test_all_stages_bit_exact(golden):
    """Every stage the reference dumped -- lookups (pos, rix, enc32), per-(strand, leaf) histograms, hdist_filt, rho,
    Brent d/v at full precision, deterministic summarize and placement -- equals the oracle bit for bit."""
    import oracle_lib as O
    p = O.default_params(want_lookups=1, want_place=1, no_filter=0)
    nplace = 0
    assert len(golden["dump"]["reads"]) == len(golden["reads"]) == 236
    for i, s in enumerate(golden["reads"]):
        o, r = golden["oracle"].query(s, p), golden["dump"]["reads"][i]
        assert r["name"] == golden["names"][i]
        for key in ("onmers", "wn", "hdist_filt", "lookups"):
            assert o[key] == r[key], (i, key)
        key_m = lambda m: (m["strand"], m["leaf_se"], m["match"], m["hdist_min"], m["rho"], m["hist"])
        assert [key_m(m) for m in o["minfo"]] == [key_m(m) for m in r["minfo"]], i
        key_s = lambda s_: (s_["leaf_se"], s_["strand"], s_["d"], s_["v"], s_["chisq"], s_["is_closest"])
        assert [key_s(x) for x in o["sel"]] == [key_s(x) for x in r["sel"]], i
        assert [tuple(q.values()) for q in o["place"]] == [tuple(q.values()) for q in r["place"]], i
        nplace += len(r["place"])
    assert nplace == 408


def test_reference_hash_order_agrees_on_untied_reads(golden):
    """The reference's own summarize_matches (pointer-hash order; 'H'/'C' lines) agrees with the deterministic rule:
    node_to_minfo always, closest whenever the minimum distance is not tied."""
    for i, r in golden["dump"]["reads"].items():
        assert sorted((h["leaf_se"], h["d"], h["v"]) for h in r["href"]) == sorted((s["leaf_se"], s["d"], s["v"]) for s in r["sel"]) or \
            _only_tie_differs(r), i
        if r["sel"]:
            dmin = min(s["d"] for s in r["sel"])
            if sum(1 for s in r["sel"] if s["d"] == dmin) == 1:
                assert r["cref"] == next(s["leaf_se"] for s in r["sel"] if s["is_closest"]), i


def _only_tie_differs(r):
    # when the closest is tied between the two strands of one leaf, the reference may keep either strand's Minfo
    a = {h["leaf_se"]: (h["d"], h["v"]) for h in r["href"]}
    b = {s["leaf_se"]: (s["d"], s["v"]) for s in r["sel"]}
    return a.keys() == b.keys() and all(a[k][0] == b[k][0] for k in a)


def test_dist_tsv_equals_reference_cli(golden):
    """`krepp dist` body printed by the reference binary == the oracle's report_distances, as sorted line sets."""
    import ctypes as C
    import oracle_lib as O
    L = O.lib()
    with open(os.path.join(SMALL, "reads.fq"), "rb") as f:
        txt = f.read()

    class Rec(C.Structure):
        _fields_ = [("name", C.c_char_p), ("seq", C.c_char_p), ("len", C.c_uint64)]
    L.ko_parse_reads.restype = C.c_int64
    L.ko_parse_reads.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(Rec))]
    L.ko_dist_tsv.restype = C.c_void_p
    L.ko_dist_tsv.argtypes = [C.c_void_p, C.POINTER(O.Params), C.POINTER(Rec), C.c_int64, C.c_int]
    recs = C.POINTER(Rec)()
    n = L.ko_parse_reads(txt, len(txt), C.byref(recs))
    assert n == 236 and recs[0].name == b"r0"
    p = O.default_params()
    ptr = L.ko_dist_tsv(golden["oracle"].h, C.byref(p), recs, n, 2)
    got = C.string_at(ptr).decode().splitlines()
    with open(os.path.join(SMALL, "ref_dist.tsv")) as f:
        ref = f.read().splitlines()
    assert sorted(got) == sorted(ref)
    assert sum(1 for l in ref if l.endswith("\tNA\tNaN")) > 0


def test_jplace_equals_reference_cli_on_untied_reads(golden):
    """Raw `krepp place` output of the reference: for reads whose closest reference is not tied, every jplace field at
    the reference's 5 decimals equals the oracle's placement; the edge-numbered tree string is identical."""
    import oracle_lib as O
    with open(os.path.join(SMALL, "ref_place.jplace")) as f:
        jp = json.load(f)
    assert jp["tree"] == golden["oracle"].jplace_tree()
    ref = {pl["n"][0]: sorted(tuple(round(x, 5) for x in row) for row in pl["p"]) for pl in jp["placements"]}
    p = O.default_params(want_place=1, no_filter=0)
    checked = 0
    for name, s in zip(golden["names"], golden["reads"]):
        o = golden["oracle"].query(s, p)
        if not o["sel"]:
            assert name not in ref
            continue
        dmin = min(x["d"] for x in o["sel"])
        if sum(1 for x in o["sel"] if x["d"] == dmin) != 1:
            continue  # tied closest: the reference's own output varies between runs
        rows = sorted((q["edge"], round(q["pendant"], 5), round(q["distal"], 5), round(-q["v"], 5), round(q["lwr"], 5), round(q["d"], 5))
                      for q in o["place"])
        if not rows:
            assert name not in ref
            continue
        assert name in ref, name
        assert len(rows) == len(ref[name]), name
        for a, b in zip(rows, ref[name]):
            assert a[0] == b[0] and all(abs(x - y) <= 1.001e-5 for x, y in zip(a[1:], b[1:])), (name, a, b)
        checked += 1
    assert checked > 100


def test_index_side_minimizers_reproduce_reference_index(golden):
    """a17: the restated RSeq::extract_mers + per-bucket sort/unique reproduces the reference-built index's inc-* and the
    enc column of cmer-* exactly (both are deterministic across builds)."""
    import ctypes as C
    import oracle_lib as O
    import synth
    L = O.lib()
    md = open(os.path.join(SMALL, "index", "metadata-m4r1-frac"), "rb").read()
    k, w, h = md[0], md[1], md[2]
    m, r = int.from_bytes(md[3:7], "little"), int.from_bytes(md[7:11], "little")
    frac, nrows = md[11], int.from_bytes(md[12:16], "little")
    geom = L.ko_geom_new(k, h, m, r, frac, md[16:16 + h])
    out, n, cap = C.POINTER(C.c_uint64)(), C.c_uint64(0), C.c_uint64(0)
    for line in open(os.path.join(SMALL, "input_map.tsv")):
        name, path = line.split()
        for _, s in synth.read_fasta(os.path.join(SMALL, path)):
            L.ko_extract_mers(geom, s.tobytes(), len(s), w, C.byref(out), C.byref(n), C.byref(cap))
    mers = np.unique(np.ctypeslib.as_array(out, shape=(n.value,)).copy())
    cmer = np.fromfile(os.path.join(SMALL, "index", "cmer-m4r1-frac"), dtype="<u4", offset=8).reshape(-1, 2)
    inc = np.fromfile(os.path.join(SMALL, "index", "inc-m4r1-frac"), dtype="<u8", offset=4)
    assert len(inc) == nrows and len(mers) == len(cmer) == inc[-1]
    rows = (mers >> np.uint64(32)).astype(np.int64)
    assert np.array_equal(np.cumsum(np.bincount(rows, minlength=nrows)), inc.astype(np.int64))
    assert np.array_equal((mers & np.uint64(0xFFFFFFFF)).astype(np.uint32), cmer[:, 0])


def test_primitives_against_python_definitions():
    import oracle_lib as O
    L = O.lib()
    rng = np.random.default_rng(0)

    def pext(x, mask):
        r, b = 0, 0
        for i in range(64):
            if (mask >> i) & 1:
                r |= ((x >> i) & 1) << b
                b += 1
        return r
    for _ in range(300):
        x, mk = int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2)), int(rng.integers(0, 2**63))
        assert L.ko_pext64(x, mk) == pext(x, mk)
    comp = {0: 3, 1: 2, 2: 1, 3: 0}
    for k in (19, 21, 27, 31, 32):
        for _ in range(50):
            codes = [int(c) for c in rng.integers(0, 4, size=k)]
            bp = 0
            for c in codes:
                bp = (bp << 2) | c
            rc = 0
            for c in reversed(codes):
                rc = (rc << 2) | comp[c]
            assert L.ko_revcomp_bp64(bp, k) == rc
            lr = L.ko_conv_bp64_lr64(bp)
            lo = sum((codes[k - 1 - p] & 1) << p for p in range(k))
            hi = sum((codes[k - 1 - p] >> 1) << p for p in range(k))
            assert lr == (hi << 32) | lo
    for _ in range(100):
        z = int(rng.integers(0, 2**32))
        assert L.ko_popcount_lr32(z) == bin((z | (z >> 16)) & 0xFFFF).count("1")
    assert L.ko_xur64_hash(0) == 0 and L.ko_xur64_hash(1) == 0xB456BCFC34C2CB2C


def test_likelihood_tables_and_known_answers():
    """Binomial tables of HDistHistLLH and three solves whose values the survey recorded from the reference
    (toy index, reads ||61435-4122 / -4949 / -317)."""
    import ctypes as C
    import oracle_lib as O
    L = O.lib()
    ck, hnk = (C.c_uint64 * 65)(), (C.c_uint64 * 17)()
    L.ko_llh_tables(11, 27, 4, ck, hnk)
    assert [ck[i] for i in range(28)] == [math.comb(27, i) for i in range(28)]
    assert [hnk[i] for i in range(5)] == [0] + [math.comb(27, i) - math.comb(16, i) for i in range(1, 5)]
    cases = [([3, 0, 0, 0, 0], 121, 0.099175417711688627, 0.048993740259441941, 11.173854718015981),
             ([6, 2, 1, 0, 0], 115, 0.097799312264855692, 0.027593291950832377, 29.627208991355744),
             ([12, 1, 0, 0, 0], 111, 0.10010962359931216, 0.0044993217180665807, 18.094698980243127)]
    for hist, uc, rho, d_ref, v_ref in cases:
        hh = (C.c_double * 17)(*hist)
        d, v, it = C.c_double(), C.c_double(), C.c_uint32()
        L.ko_brent(11, 27, 4, hh, uc, rho, C.byref(d), C.byref(v), C.byref(it))
        assert d.value == d_ref and v.value == v_ref and 5 <= it.value <= 30
