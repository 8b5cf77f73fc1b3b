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
