"""CPU: the restatement of `krepp seek` (oracle ko_sketch_load / ko_seek_read) pinned on the UNMODIFIED reference -- sketches
built by `krepp sketch`, queried by `krepp seek`, every printed row equal (ref src/seek.cpp:22-127, src/sketch.cpp:3-39)."""
import os

import pytest

from conftest import needs_ref
from sketches import SKETCHES, SMALL, build_sketch, ref_seek
from test_gpu_parity import fastq_reads

pytestmark = needs_ref


@pytest.mark.parametrize("label,genome,args", SKETCHES, ids=[s[0] for s in SKETCHES])
def test_oracle_seek_equals_the_reference(label, genome, args, tmp_path_factory):
    import oracle_lib as O
    path = build_sketch(label, genome, args, tmp_path_factory.getbasetemp())
    names, reads = fastq_reads(os.path.join(SMALL, "reads.fq"))
    sk = O.OracleSketch(path)
    for th in (4, 2) if label == "default" else (4,):
        ref = ref_seek(path, os.path.join(SMALL, "reads.fq"), th)
        assert ref[0].startswith("r")  # the reference builds a header (src/krepp.cpp:305-309) and never writes it: rows only
        mine = [sk.tsv_row(n, r, th) for n, r in zip(names, reads)]
        assert sorted(mine) == sorted(ref), (label, th)
        found = sum(1 for x in mine if not x.endswith("NaN"))
        assert found > 10 and found < len(mine), (label, found)


def test_sketch_loads_as_a_one_reference_image_and_rows_format(tmp_path_factory):
    """Host side without a GPU: the sketch file as an index image (geometry, one leaf named after the file, bucket statistics),
    a truncated file refused, and the row formatter (std::fixed precision 5, NaN for a read without a match)."""
    import numpy as np
    import krepp_b200
    import oracle_lib as O
    from krepp_b200 import capi
    label, genome, args = SKETCHES[2]
    path = build_sketch(label, genome, args, tmp_path_factory.getbasetemp())
    ix = krepp_b200.Index(path, capi.DEVICE_NONE)
    assert (ix.info.k, ix.info.w, ix.info.h, ix.info.m, ix.info.r, ix.info.frac) == (21, 21, 7, 3, 1, 1)
    assert ix.info.nleaves == 1 and ix.info.nnodes == 1 and ix.info.nsubsets == 2 and ix.info.nkmers > 1000
    assert ix.jplace_tree() == os.path.basename(path) + "{0};" and O.OracleSketch(path).k == 21
    ix.close()
    cut = os.path.join(os.path.dirname(path), "cut.skc")
    with open(path, "rb") as f, open(cut, "wb") as g:
        g.write(f.read()[:-9])
    with pytest.raises(capi.KreppError, match="Failed to read the sketch file!"):
        krepp_b200.Index(cut, capi.DEVICE_NONE)
    txt = capi.format_seek(np.array([0.015994, float("nan"), 0.5, 1e-10, 0.123455]), ["a", "b", "c d", "e", "f"])
    assert txt == "a\t0.01599\nb\tNaN\nc d\t0.50000\ne\t0.00000\nf\t" + "%.5f" % 0.123455 + "\n"


def fasta_seqs(path):
    seqs, cur = [], []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if cur:
                    seqs.append("".join(cur).encode())
                cur = []
            else:
                cur.append(line.strip())
    if cur:
        seqs.append("".join(cur).encode())
    return seqs


@pytest.mark.parametrize("label,genome,args", SKETCHES, ids=[s[0] for s in SKETCHES])
def test_oracle_sketch_equals_the_reference_file(label, genome, args, tmp_path_factory):
    """`krepp sketch` restated (extract_mers walk, per-row sort + unique, HyperLogLog rho) against the file the reference wrote:
    the table entry for entry, rho bit for bit (ref src/krepp.cpp:110-128, src/rqseq.cpp:51-144, src/hyperloglog.hpp:58-140)."""
    import numpy as np
    import oracle_lib as O
    path = build_sketch(label, genome, args, tmp_path_factory.getbasetemp())
    meta = O.read_sketch_file(path)
    keys, rho = O.oracle_sketch_table(meta, fasta_seqs(os.path.join(SMALL, "genomes", genome + ".fna")))
    assert len(keys) == len(meta["enc"]) > 500
    assert np.array_equal((keys & 0xFFFFFFFF).astype(np.uint32), meta["enc"])
    inc = np.cumsum(np.bincount((keys >> 32).astype(np.int64), minlength=meta["nrows"])).astype(np.uint64)
    assert np.array_equal(inc, meta["inc"])
    assert rho == meta["rho"], (rho, meta["rho"])
