"""GPU: long reads cut into segments (SegArgs / segment_combine_kernel, SURVEY.md 8 "missing 5": a contig is one query in the
reference, src/rqseq.cpp:189, src/query.cpp:46).  The segments' histograms summed up again must be the read's own bit for bit:
against the oracle stage by stage, against the uncut path, in both forms of the match step, for dist and place, with segment
lengths from 32 windows (every golden read is cut) to the default (only the contig-sized reads are)."""
import os
import subprocess

import numpy as np
import pytest

import conftest
from conftest import needs_ref
from test_gpu_parity import fastq_reads
from variants import SMALL

pytestmark = [pytest.mark.gpu]
EXE = os.path.join(conftest.ROOT, "krepp_b200", "_build", "krepp_b200")
IDX = os.path.join(SMALL, "index")


def genome(name):
    with open(os.path.join(SMALL, "genomes", name + ".fna")) as f:
        return "".join(l.strip() for l in f if not l.startswith(">")).encode()


def long_reads():
    """Contig-sized queries: two whole genomes, one with a stretch of N (runs reset), a chimera of two genomes, a read just over
    and one just under the cutting threshold, and short reads in between so that cut and uncut reads share a batch."""
    _, short = fastq_reads(os.path.join(SMALL, "reads.fq"))
    g0, g3, g6 = genome("G000000"), genome("G000003"), genome("G000006")
    withn = g3[:7000] + b"N" * 40 + g3[7040:15000] + b"NNN" + g3[15003:]
    return [short[0], g0, short[1], withn, g6[:9000] + g0[3000:12000], short[2], g6[:2 * 1024 + 27 - 1], g6[:2 * 1024 + 27], short[3]]


def run(seg, reads, pipeline, monkeypatch, **kw):
    import krepp_b200
    monkeypatch.setenv("KREPP_PIPELINE", pipeline)
    monkeypatch.setenv("KREPP_SEGMENT_WINDOWS", str(seg))
    g = krepp_b200.Index(IDX, 0)
    b = krepp_b200.IBatch(g, reads, **kw)
    res = b.results()
    out = {k: np.array(res[k], copy=True) for k in ("reads", "records", "hist", "placements")}
    out["stages"] = [n for n, _ in b.stage_times()]
    b.close()
    g.close()
    return out


def canon(res):
    """Per read: its summary without the positions of its rows (rows are handed out by atomics, so their place in the batch
    changes from run to run), and the rows themselves, bytes."""
    out = []
    rd = res["reads"]
    for i in range(len(rd)):
        s = rd[i]
        rb, rc, pb, pc = int(s["rec_begin"]), int(s["rec_count"]), int(s["place_begin"]), int(s["place_count"])
        head = (int(s["onmers"]), s["wn"].tolist(), s["hdist_filt"].tolist(), rc, pc, int(s["closest"]) - rb if s["closest"] >= 0 else -1, int(s["n_selected"]))
        out.append((head, res["records"][rb:rb + rc].tobytes(), res["hist"][rb:rb + rc].tobytes(), res["placements"][pb:pb + pc].tobytes()))
    return out


def same(a, b):
    ca, cb = canon(a), canon(b)
    assert len(ca) == len(cb)
    for i, (x, y) in enumerate(zip(ca, cb)):
        assert x[0] == y[0], (i, x[0], y[0])
        assert x[1:] == y[1:], (i, "rows differ")


@pytest.mark.parametrize("pipeline", ["sorted", "fused"])
def test_cut_reads_equal_uncut_reads_bit_for_bit(pipeline, monkeypatch):
    reads = long_reads()
    for kw in (dict(), dict(place=True, no_filter=False)):
        whole = run(0, reads, pipeline, monkeypatch, **kw)
        assert "segment_combine_kernel" not in whole["stages"]
        assert int(whole["reads"]["rec_count"][1]) > 2 and len(whole["records"]) > 20
        for seg in (1024, 100, 32):
            cut = run(seg, reads, pipeline, monkeypatch, **kw)
            assert "segment_combine_kernel" in cut["stages"], seg
            same(whole, cut)


@pytest.mark.parametrize("pipeline", ["sorted", "fused"])
def test_cut_reads_equal_oracle(pipeline, monkeypatch):
    """Every golden read cut into 32-window segments, and the contig-sized reads at the default length: stages against the oracle
    (tap 2: every (strand, reference) pair before the gate, so the segments' sums are compared for all of them)."""
    import krepp_b200
    import oracle_lib as O
    from gpu_common import run_and_compare
    monkeypatch.setenv("KREPP_PIPELINE", pipeline)
    o = O.OracleIndex(IDX)
    _, short = fastq_reads(os.path.join(SMALL, "reads.fq"))
    for seg, reads in ((32, short[:200]), (1024, long_reads())):
        monkeypatch.setenv("KREPP_SEGMENT_WINDOWS", str(seg))
        g = krepp_b200.Index(IDX, 0)
        st = run_and_compare(IDX, reads, o, g, check_lookups=False)
        assert st["solves"] > 5, st
        st = run_and_compare(IDX, reads, o, g, check_lookups=False, place=True, no_filter=False)
        assert st["placements"] > 2, st
        g.close()


@needs_ref
def test_contigs_through_the_executable_equal_the_reference(tmp_path):
    """Whole genomes as queries (FASTA) through `krepp_b200 dist` against the reference CLI."""
    fa = tmp_path / "contigs.fa"
    with open(fa, "w") as f:
        for i, s in enumerate(long_reads()):
            f.write(f">c{i}\n{s.decode()}\n")
    ref = subprocess.run([os.path.join(conftest.REF_DIR, "krepp"), "dist", "-i", IDX, "-q", str(fa)], capture_output=True, text=True, check=True).stdout.splitlines()[2:]
    mine = subprocess.run([EXE, "dist", "-i", IDX, "-q", str(fa)], capture_output=True, text=True, check=True).stdout.splitlines()[2:]
    assert sorted(mine) == sorted(ref) and len(mine) > 20
