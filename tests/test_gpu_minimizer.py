"""GPU test of the index-side k-mer kernel (krepp_b200/csrc/minimizer.cu): window minimizers, LSH
residue filter, rows and residual encodings of whole genomes, sorted and made unique per row -- bit-exact against the oracle's
restatement of RSeq::extract_mers (ref src/rqseq.cpp:51-144, pinned against the reference-built toy index) on sequences with N
runs, lower case, short contigs and every end-of-sequence case, and against the reference-built index itself (inc-* and the enc
column of cmer-*, which are deterministic across builds: SURVEY.md section 0 fact 4)."""
import ctypes as C
import os

import numpy as np
import pytest

import conftest

pytestmark = [pytest.mark.gpu]
SMALL = os.path.join(conftest.GOLDEN_DIR, "small")


def oracle_keys(index_dir, seqs):
    import oracle_lib as O
    L = O.lib()
    sfx = [f[8:] for f in os.listdir(index_dir) if f.startswith("metadata-") and "." not in f][0]
    md = open(os.path.join(index_dir, "metadata" + sfx), "rb").read()
    k, w, h = md[0], md[1], md[2]
    m, r, frac = int.from_bytes(md[3:7], "little"), int.from_bytes(md[7:11], "little"), md[11]
    geom = L.ko_geom_new(k, h, m, r, frac, md[16:16 + h])
    out, n, cap = C.POINTER(C.c_uint64)(), C.c_uint64(0), C.c_uint64(0)
    for s in seqs:
        L.ko_extract_mers(geom, s, len(s), w, C.byref(out), C.byref(n), C.byref(cap))
    return np.unique(np.ctypeslib.as_array(out, shape=(n.value,)).copy()) if n.value else np.zeros(0, np.uint64)


def test_minimizer_kernel_reproduces_the_reference_built_index():
    import krepp_b200
    import synth
    g = krepp_b200.Index(os.path.join(SMALL, "index"), 0)
    per_genome = []
    for line in open(os.path.join(SMALL, "input_map.tsv")):
        name, path = line.split()
        seqs = [s.tobytes() for _, s in synth.read_fasta(os.path.join(SMALL, path))]
        keys = g.extract_mers(seqs)
        assert np.array_equal(keys, oracle_keys(os.path.join(SMALL, "index"), seqs)), name   # one leaf table, bit for bit
        assert len(keys) > 1000 and np.all(keys[1:] > keys[:-1])
        per_genome.append(keys)
    mers = np.unique(np.concatenate(per_genome))
    cmer = np.fromfile(os.path.join(SMALL, "index", "cmer-m4r1-frac"), dtype="<u4", offset=8).reshape(-1, 2)
    inc = np.fromfile(os.path.join(SMALL, "index", "inc-m4r1-frac"), dtype="<u8", offset=4)
    assert len(mers) == len(cmer) == inc[-1]
    rows = (mers >> np.uint64(32)).astype(np.int64)
    assert np.array_equal(np.cumsum(np.bincount(rows, minlength=len(inc))), inc.astype(np.int64))
    assert np.array_equal((mers & np.uint64(0xFFFFFFFF)).astype(np.uint32), cmer[:, 0])


@pytest.mark.parametrize("index_dir", ["small", "toy"])
def test_minimizer_kernel_edge_cases(index_dir):
    """N runs, lower case, contigs shorter than w / exactly w, every end-of-sequence case of ref src/rqseq.cpp:112-116 (last valid
    run shorter than w with earlier k-mers in the ring, with too few k-mers in the whole sequence -> zero slots), tile borders."""
    import krepp_b200
    if index_dir == "toy":
        d = os.path.join(conftest.TOY_DIR, "index_toy")
        if not os.path.isdir(d):
            pytest.skip("oracle/_ref/toy not built")
    else:
        d = os.path.join(SMALL, "index")
    g = krepp_b200.Index(d, 0)
    k, w = g.info.k, g.info.w
    rng = np.random.default_rng(9)

    def rnd(n):
        return bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=n))
    cases = [[rnd(5000)], [rnd(w - 1)], [rnd(w)], [rnd(w + 1)], [rnd(3000).lower()],
             [rnd(2000) + b"N" * 7 + rnd(k)], [rnd(2000) + b"N" + rnd(k + 3)], [rnd(2000) + b"NNN" + rnd(w - 1)], [rnd(2000) + b"N" + rnd(w)],
             [rnd(k + 2) + b"N" + rnd(k + 1)], [b"N" * 40 + rnd(k)], [rnd(10) + b"N" + rnd(w - 2)], [rnd(k) + b"N" * 100],
             [rnd(700) + b"N" + rnd(k - 1)], [b"A" * 300], [b"ACGT" * 100 + b"R" + b"ACGT" * 8],
             [rnd(120 + w - 1)], [rnd(121 + w - 1)], [rnd(119 + w - 1)], [rnd(240 + w - 1)], [rnd(241 + w - 1)],
             [rnd(int(rng.integers(1, 900))) for _ in range(60)]]
    for _ in range(25):  # random sequences sprinkled with N
        s = bytearray(rnd(int(rng.integers(w, 4000))))
        for _ in range(int(rng.integers(0, 12))):
            at = int(rng.integers(0, len(s)))
            s[at:at + int(rng.integers(1, 30))] = b"N" * min(int(rng.integers(1, 30)), len(s) - at)
        cases.append([bytes(s)])
    for i, seqs in enumerate(cases):
        want = oracle_keys(d, seqs)
        got = g.extract_mers(seqs)
        assert np.array_equal(got, want), (i, [len(s) for s in seqs], len(got), len(want))
    big = [rnd(300_000), rnd(50_000) + b"N" * 1000 + rnd(120_000), rnd(w + 5)]
    assert np.array_equal(g.extract_mers(big), oracle_keys(d, big))
