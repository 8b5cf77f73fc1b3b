"""CPU: the oracle against the unmodified reference on index geometries other than the default (tests/variants.py): stage
dump (lookups, histograms, solves, placements) of `ref_dump` and the `krepp dist` TSV.  Pins the oracle for the rows the
GPU variant test (tests/test_gpu_variants.py) then checks the CUDA path against."""
import os
import subprocess

import pytest

from conftest import REF_DIR, needs_ref
from test_gpu_parity import fastq_reads
from variants import SMALL, TREELESS, VARIANTS, build

pytestmark = [needs_ref]


@pytest.mark.parametrize("label,args", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_oracle_equals_reference_on_variant(label, args, tmp_path_factory):
    import oracle_lib as O
    idx = build(label, args, tmp_path_factory.getbasetemp())
    fq = os.path.join(SMALL, "reads.fq")
    names, reads = fastq_reads(fq)
    dump = O.parse_ref_dump(subprocess.run([os.path.join(REF_DIR, "ref_dump"), idx, fq, "--lookups", "--place"], capture_output=True, text=True,
                                           check=True).stdout)
    ix = O.OracleIndex(idx)
    p = O.default_params(want_lookups=1, want_place=1, no_filter=0)
    nrec = 0
    for i, s in enumerate(reads):
        o, r = ix.query(s, p), dump["reads"][i]
        for key in ("onmers", "wn", "hdist_filt", "lookups"):
            assert o[key] == r[key], (label, i, key)
        key_m = lambda m: (m["strand"], m["leaf_se"], m["match"], m["hdist_min"], m["rho"], m["hist"])
        assert [key_m(m) for m in o["minfo"]] == [key_m(m) for m in r["minfo"]], (label, i)
        key_s = lambda s_: (s_["leaf_se"], s_["strand"], s_["d"], s_["v"], s_["chisq"], s_["is_closest"])
        assert [key_s(x) for x in o["sel"]] == [key_s(x) for x in r["sel"]], (label, i)
        assert [tuple(q.values()) for q in o["place"]] == [tuple(q.values()) for q in r["place"]], (label, i)
        nrec += len(o["minfo"])
    assert nrec > 40, (label, nrec)
    ref = subprocess.run([os.path.join(REF_DIR, "krepp"), "dist", "-i", idx, "-q", fq], capture_output=True, text=True, check=True).stdout.splitlines()[2:]
    ora = subprocess.run([os.path.join(os.path.dirname(REF_DIR), "_build", "krepp_oracle"), "dist", idx, fq], capture_output=True, text=True,
                         check=True).stdout.splitlines()[1:]
    assert sorted(ref) == sorted(ora), label


def test_treeless_index_dist_equals_reference(tmp_path_factory):
    """No tree-* file: the tree is the balanced one the reference generates over reflist-*.  The oracle's and the C++
    loader's restatement of it must give the reference's leaf numbering (else every reference name below would be wrong)."""
    import krepp_b200
    import oracle_lib as O
    idx = build(*TREELESS, tmp_path_factory.getbasetemp(), with_tree=False)
    assert os.path.exists(os.path.join(idx, "reflist-m4r1-frac")) and not os.path.exists(os.path.join(idx, "tree-m4r1-frac"))
    fq = os.path.join(SMALL, "reads.fq")
    ref = subprocess.run([os.path.join(REF_DIR, "krepp"), "dist", "-i", idx, "-q", fq], capture_output=True, text=True, check=True).stdout.splitlines()[2:]
    ora = subprocess.run([os.path.join(os.path.dirname(REF_DIR), "_build", "krepp_oracle"), "dist", idx, fq], capture_output=True, text=True,
                         check=True).stdout.splitlines()[1:]
    assert len(ref) > 300 and sorted(ref) == sorted(ora)
    ix, o = krepp_b200.Index(idx, device=-1), O.OracleIndex(idx)   # the C++ loader builds the same tree as the oracle
    assert ix.info.nnodes == 15 and ix.info.nleaves == 8
    for se in range(1, ix.info.nnodes + 1):
        assert ix.node_name(se) == o.name(se), se
    place = subprocess.run([os.path.join(REF_DIR, "krepp"), "place", "-i", idx, "-q", fq], capture_output=True, text=True)
    assert place.returncode != 0 and "lacks a tree" in place.stderr   # what krepp_batch_create answers for place on this handle


def test_partial_library_directory_oracle_pinned_loader_merges(tmp_path_factory):
    """A directory holding several partial libraries (three `krepp index --no-frac` runs with r = 0, 2, 3 of m = 4 into one
    -o directory; every partial has its own table, colour record and rho).  The oracle restates the per-residue dispatch
    (ref src/index.cpp:144-168) and is pinned here against the reference's `dist`; the C++ loader merges the partials into one
    image (tables one after the other, colour ids shifted per partial), checked here for its sizes and on the GPU against the
    oracle (tests/test_gpu_variants.py)."""
    import numpy as np
    import krepp_b200
    from variants import build_partials
    idx, fq = build_partials(tmp_path_factory.getbasetemp()), os.path.join(SMALL, "reads.fq")
    ref = subprocess.run([os.path.join(REF_DIR, "krepp"), "dist", "-i", idx, "-q", fq], capture_output=True, text=True, check=True).stdout.splitlines()[2:]
    ora = subprocess.run([os.path.join(os.path.dirname(REF_DIR), "_build", "krepp_oracle"), "dist", idx, fq], capture_output=True, text=True,
                         check=True).stdout.splitlines()[1:]
    assert len(ref) > 1000 and sorted(ref) == sorted(ora)
    ix = krepp_b200.Index(idx, device=-1)
    sizes = {}
    for sfx in ("-m4r0-no_frac", "-m4r2-no_frac", "-m4r3-no_frac"):
        nk = int(np.fromfile(os.path.join(idx, "cmer" + sfx), dtype="<u8", count=1)[0])
        nr = int(np.fromfile(os.path.join(idx, "inc" + sfx), dtype="<u4", count=1)[0])
        ns = int(np.fromfile(os.path.join(idx, "crecord" + sfx), dtype="<u4", count=2)[1])
        sizes[sfx] = (nk, nr, ns)
    assert ix.info.nkmers == sum(v[0] for v in sizes.values()) and ix.info.nrows == sum(v[1] for v in sizes.values())
    assert ix.info.nsubsets == (ix.info.nnodes + 1) + sum(v[2] - (ix.info.nnodes + 1) for v in sizes.values())
    cs = ix.host_checksums()   # sum of the bucket ends: every partial's ends, moved up by the entries of the partials before it
    want, base = 0, 0
    for sfx in sorted(sizes):
        inc = np.fromfile(os.path.join(idx, "inc" + sfx), dtype="<u8", offset=4)
        want += int(inc.sum()) + base * len(inc)
        base += sizes[sfx][0]
    assert cs[1] == want % (1 << 64)
    # overlapping partials (two runs that both hold residue 0) are refused
    from krepp_b200.capi import KreppError
    both = build_partials(tmp_path_factory.getbasetemp(), rs=("0", "1"), frac=True)
    with pytest.raises(KreppError, match="overlap"):
        krepp_b200.Index(both, device=-1)
