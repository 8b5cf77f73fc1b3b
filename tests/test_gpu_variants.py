"""GPU: the CUDA path against the oracle on index geometries other than the default (tests/variants.py; the oracle itself is
pinned on them against the reference in tests/test_variants_cpu.py): k from 19 to 31 (7- and 8-table LUTs, residual
encodings of 14 to 16 positions), m = 1, 2, 3, 4, 5 with frac and no-frac row addressing, dist and place, both forms of the
match step, and the reference's own `krepp dist` output through the C++ formatters."""
import os
import subprocess

import pytest

from conftest import REF_DIR, needs_ref
from test_gpu_parity import fastq_reads
from variants import SMALL, TREELESS, VARIANTS, build

pytestmark = [pytest.mark.gpu, needs_ref]


@pytest.mark.parametrize("pipeline", ["fused", "sorted"])
@pytest.mark.parametrize("label,args", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_gpu_equals_oracle_on_variant(label, args, pipeline, tmp_path_factory, monkeypatch):
    import krepp_b200
    import oracle_lib as O
    from gpu_common import run_and_compare
    monkeypatch.setenv("KREPP_PIPELINE", pipeline)
    idx = build(label, args, tmp_path_factory.getbasetemp())
    names, reads = fastq_reads(os.path.join(SMALL, "reads.fq"))
    o, g = O.OracleIndex(idx), krepp_b200.Index(idx, 0)
    st = run_and_compare(idx, reads, o, g)
    assert st["solves"] > 20, (label, st)
    st = run_and_compare(idx, reads, o, g, check_lookups=False, place=True, no_filter=False)
    assert st["placements"] > 10, (label, st)
    b = krepp_b200.IBatch(g, reads, names=names)
    ref = subprocess.run([os.path.join(REF_DIR, "krepp"), "dist", "-i", idx, "-q", os.path.join(SMALL, "reads.fq")], capture_output=True, text=True,
                         check=True).stdout.splitlines()[2:]
    assert sorted(b.estimate_distances().splitlines()) == sorted(ref), label
    b.close()
    g.close()


def test_gpu_treeless_index(tmp_path_factory):
    """dist on an index built without a backbone tree (tree generated from reflist-*): stages against the oracle, TSV against
    the reference; place is refused with the reference's message."""
    import krepp_b200
    import oracle_lib as O
    from gpu_common import run_and_compare
    from krepp_b200.capi import KreppError
    idx = build(*TREELESS, tmp_path_factory.getbasetemp(), with_tree=False)
    names, reads = fastq_reads(os.path.join(SMALL, "reads.fq"))
    o, g = O.OracleIndex(idx), krepp_b200.Index(idx, 0)
    st = run_and_compare(idx, reads, o, g)
    assert st["solves"] > 500
    b = krepp_b200.IBatch(g, reads, names=names)
    ref = subprocess.run([os.path.join(REF_DIR, "krepp"), "dist", "-i", idx, "-q", os.path.join(SMALL, "reads.fq")], capture_output=True, text=True,
                         check=True).stdout.splitlines()[2:]
    assert sorted(b.estimate_distances().splitlines()) == sorted(ref)
    b.close()
    with pytest.raises(KreppError, match="lacks a tree"):
        krepp_b200.IBatch(g, reads[:4], place=True, no_filter=False)
    g.close()


@pytest.mark.parametrize("pipeline", ["fused", "sorted"])
def test_gpu_partial_library_directory(pipeline, tmp_path_factory, monkeypatch):
    """Three partial libraries in one directory (residues 0, 2, 3 of m = 4, each with its own table and colour record; ref
    src/krepp.cpp:66-108, src/index.cpp:144-168): the merged image against the oracle on every stage, dist and place, the TSV
    against the reference, and the merged table split into bucket-range shards (logical ranks on one GPU)."""
    import krepp_b200
    import oracle_lib as O
    from gpu_common import run_and_compare
    from variants import build_partials
    monkeypatch.setenv("KREPP_PIPELINE", pipeline)
    idx = build_partials(tmp_path_factory.getbasetemp())
    names, reads = fastq_reads(os.path.join(SMALL, "reads.fq"))
    o, g = O.OracleIndex(idx), krepp_b200.Index(idx, 0)
    st = run_and_compare(idx, reads, o, g)
    assert st["solves"] > 500, st
    st = run_and_compare(idx, reads, o, g, check_lookups=False, place=True, no_filter=False)
    assert st["placements"] > 100, st
    b = krepp_b200.IBatch(g, reads, names=names)
    ref = subprocess.run([os.path.join(REF_DIR, "krepp"), "dist", "-i", idx, "-q", os.path.join(SMALL, "reads.fq")], capture_output=True, text=True,
                         check=True).stdout.splitlines()[2:]
    mine = b.estimate_distances().splitlines()
    assert sorted(mine) == sorted(ref)
    b.close()
    g.close()
    if pipeline == "sorted":
        exe = os.path.join(os.path.dirname(REF_DIR), "..", "krepp_b200", "_build", "krepp_b200")
        one = subprocess.run([exe, "dist", "-i", idx, "-q", os.path.join(SMALL, "reads.fq")], capture_output=True, text=True, check=True).stdout.splitlines()[2:]
        sh = subprocess.run([exe, "dist", "-i", idx, "-q", os.path.join(SMALL, "reads.fq"), "--shard-index", "--devices", "0,0,0"], capture_output=True, text=True,
                            check=True).stdout.splitlines()[2:]
        assert one == sh and sorted(one) == sorted(ref)
