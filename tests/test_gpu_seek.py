"""GPU: `krepp seek` -- a batch of reads against the sketch of one genome.  The CUDA path (the same match /
resolve / solve chain as `dist`, on a one-reference image of the sketch, plus seek_kernel) against the oracle's per-strand
histograms and distances, and the executable's output against the UNMODIFIED reference's `krepp seek` on sketches the reference's
`krepp sketch` built."""
import math
import os
import subprocess

import numpy as np
import pytest

import conftest
from conftest import needs_ref
from sketches import SKETCHES, SMALL, build_sketch, ref_seek
from test_gpu_parity import fastq_reads

pytestmark = [pytest.mark.gpu, needs_ref]
EXE = os.path.join(conftest.ROOT, "krepp_b200", "_build", "krepp_b200")


@pytest.mark.parametrize("pipeline", ["fused", "sorted"])
@pytest.mark.parametrize("label,genome,args", SKETCHES, ids=[s[0] for s in SKETCHES])
def test_seek_equals_oracle_and_reference(label, genome, args, pipeline, tmp_path_factory, monkeypatch):
    import krepp_b200
    import oracle_lib as O
    monkeypatch.setenv("KREPP_PIPELINE", pipeline)
    path = build_sketch(label, genome, args, tmp_path_factory.getbasetemp())
    names, reads = fastq_reads(os.path.join(SMALL, "reads.fq"))
    reads = reads + [b"ACGT" * 3, b"", b"N" * 200, reads[2][:40] + b"N" + reads[2][40:]]      # shorter than k, empty, no valid k-mer, a run broken by N
    names = names + ["short", "empty", "allN", "brokenN"]
    sk, g = O.OracleSketch(path), krepp_b200.Index(path, 0)
    assert g.info.nleaves == 1 and g.info.k == sk.k
    for th in (4, 2) if label == "default" else (4,):
        b = krepp_b200.IBatch(g, reads, names=names, hdist_th=th)
        b.set_output(seek=True)
        res = b.results()
        d, rd, rec, hist = res["seek_dist"], res["reads"], res["records"], res["hist"]
        found = 0
        for i, seq in enumerate(reads):
            o = sk.seek(seq, th)
            assert int(rd[i]["onmers"]) == o["onmers"], (label, i)
            mine = {int(r["strand"]): (hist[rd[i]["rec_begin"] + j][:th + 1].tolist(), float(r["d_llh"])) for j, r in
                    enumerate(rec[rd[i]["rec_begin"]:rd[i]["rec_begin"] + rd[i]["rec_count"]])}
            for st in range(2):
                if o["match"][st]:
                    assert mine[st][0] == o["hist"][st], (label, th, i, st)                     # bit-exact histograms
                    assert abs(mine[st][1] - o["d"][st]) <= 1e-5 * max(1.0, abs(o["d"][st])), (label, th, i, st)
                else:
                    assert st not in mine
            if o["found"]:
                found += 1
                assert abs(d[i] - o["dist"]) <= 1e-5 * max(1.0, abs(o["dist"])), (label, th, i, d[i], o["dist"])   # fp64, tolerance 1e-5 (north_star)
            else:
                assert math.isnan(d[i]), (label, th, i)
        assert found > 10
        ref = ref_seek(path, os.path.join(SMALL, "reads.fq"), th)
        assert sorted(b.seek_sequences().splitlines()[:len(ref)]) == sorted(ref), (label, th)
        b.close()
    with pytest.raises(krepp_b200.capi.KreppError, match="lacks a tree"):
        krepp_b200.IBatch(g, reads[:4], place=True, no_filter=False)
    g.close()


def test_seek_executable_equals_the_reference(tmp_path_factory, tmp_path):
    q = os.path.join(SMALL, "reads.fq")
    for label, genome, args in SKETCHES[:3]:
        path = build_sketch(label, genome, args, tmp_path_factory.getbasetemp())
        for extra in ([], ["--hdist-th", "3"], ["--num-threads", "3", "--batch-reads", "64"]):
            th = extra[1] if extra[:1] == ["--hdist-th"] else "4"
            mine = subprocess.run([EXE, "seek", "-i", path, "-q", q, *extra], capture_output=True, text=True, check=True).stdout.splitlines()
            ref = ref_seek(path, q, int(th))
            assert sorted(mine) == sorted(ref) and mine == sorted(mine, key=lambda l: int(l.split("\t")[0][1:])), (label, extra)   # rows only, input order
    out = tmp_path / "o.tsv"
    subprocess.run([EXE, "seek", "--sketch-path", path, "-q", q, "-o", str(out)], check=True, capture_output=True)
    assert sorted(out.read_text().splitlines()) == sorted(ref_seek(path, q))
    r = subprocess.run([EXE, "seek", "-i", os.path.join(SMALL, "index"), "-q", q], capture_output=True, text=True)
    assert r.returncode != 0 and "--sketch-path: File does not exist" in r.stderr
    r = subprocess.run([EXE, "dist", "-i", path, "-q", q], capture_output=True, text=True)
    assert r.returncode != 0 and "Directory does not exist" in r.stderr


def test_seek_output_needs_a_sketch_handle():
    import krepp_b200
    g = krepp_b200.Index(os.path.join(SMALL, "index"), 0)
    b = krepp_b200.IBatch(g, [b"ACGT" * 40])
    with pytest.raises(krepp_b200.capi.KreppError, match="sketch handle"):
        b.set_output(seek=True)
    b.close()
    g.close()
