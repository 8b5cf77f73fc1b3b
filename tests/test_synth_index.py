"""tools/synth_index (the on-box generator of the config 3-5 workloads) against the reference's own `krepp index`:
same genomes + same tree => identical metadata, identical inc-* and identical enc column of cmer-* (the deterministic
parts of an index), and -- colours being reproducible only up to relabelling -- the
reference's `krepp dist` prints the same distances on either index once both carry the same rho values (the generator
uses an exact minimizer/k-mer ratio where the reference uses a HyperLogLog estimate)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import conftest
from conftest import ROOT, needs_ref

TOOL_SRC = os.path.join(ROOT, "tools", "synth_index.cpp")
TOOL = os.path.join(ROOT, "tools", "_build", "synth_index")


def build_tool():
    os.makedirs(os.path.dirname(TOOL), exist_ok=True)
    if not os.path.exists(TOOL) or os.path.getmtime(TOOL) < os.path.getmtime(TOOL_SRC):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O3", "-fopenmp", "-Wall", "-o", TOOL, TOOL_SRC], check=True)
    return TOOL


def test_generator_builds_a_loadable_index(tmp_path):
    """No reference needed: the generated directory loads through the C-ABI host loader and through the oracle, and the
    two agree on geometry, tree and names; the reads file has the promised shape."""
    import krepp_b200
    import oracle_lib as O
    out = tmp_path / "w"
    subprocess.run([build_tool(), "--out", str(out), "--genomes", "9", "--length", "60000", "--reads", "500", "--fastq-reads", "500", "--seed", "5"],
                   check=True, capture_output=True)
    ix, o = krepp_b200.Index(str(out / "index"), device=-1), O.OracleIndex(str(out / "index"))
    assert (ix.info.k, ix.info.h, ix.info.m, ix.info.nnodes, ix.info.nleaves) == (27, 11, 4, 17, 9)
    assert ix.jplace_tree() == o.jplace_tree()
    assert os.path.getsize(out / "reads.u8") == 500 * 150
    names, reads, _ = krepp_b200.Reader(str(out / "reads.fq")).next_batch()
    assert len(reads) == 500 and reads[7] == open(out / "reads.u8", "rb").read()[7 * 150:8 * 150]
    # reads come from the genomes: nearly all of them hit something
    hit = sum(1 for s in reads[:200] if o.query(s)["sel"])
    assert hit > 150


@needs_ref
def test_generator_matches_reference_krepp_index(tmp_path):
    out = tmp_path / "w"
    subprocess.run([build_tool(), "--out", str(out), "--genomes", "12", "--length", "100000", "--reads", "3000", "--fastq-reads", "3000", "--fasta", "--seed", "3"],
                   check=True, capture_output=True)
    krepp = os.path.join(conftest.REF_DIR, "krepp")
    subprocess.run([krepp, "--num-threads", "4", "index", "-h", "11", "-k", "27", "-w", "35", "-o", "ref_index", "-i", "input_map.tsv", "-t", "tree.nwk"],
                   cwd=out, check=True, capture_output=True)
    sfx = "-m4r1-frac"
    mine, ref = out / "index", out / "ref_index"
    assert (mine / ("metadata" + sfx)).read_bytes() == (ref / ("metadata" + sfx)).read_bytes()
    assert (mine / ("inc" + sfx)).read_bytes() == (ref / ("inc" + sfx)).read_bytes()
    a = np.fromfile(mine / ("cmer" + sfx), dtype="<u4", offset=8).reshape(-1, 2)
    b = np.fromfile(ref / ("cmer" + sfx), dtype="<u4", offset=8).reshape(-1, 2)
    assert np.array_equal(a[:, 0], b[:, 0]) and len(a) > 50000
    # same rho on both sides, then the reference binary itself must not see a difference
    ha, hb = np.fromfile(mine / ("crecord" + sfx), dtype="<u4", count=2), np.fromfile(ref / ("crecord" + sfx), dtype="<u4", count=2)
    assert ha[0] == hb[0]
    rho_ref = np.fromfile(ref / ("crecord" + sfx), dtype="<f8", offset=8 + 8 * int(hb[1]), count=int(hb[0]))
    rho_mine = np.fromfile(mine / ("crecord" + sfx), dtype="<f8", offset=8 + 8 * int(ha[1]), count=int(ha[0]))
    assert np.all((rho_ref > 0) == (rho_mine > 0)) and np.allclose(rho_mine[rho_ref > 0], rho_ref[rho_ref > 0], rtol=0.1)
    buf = bytearray((mine / ("crecord" + sfx)).read_bytes())
    buf[8 + 8 * int(ha[1]):] = rho_ref.tobytes()
    (mine / ("crecord" + sfx)).write_bytes(bytes(buf))
    outs = []
    for d in ("index", "ref_index"):
        r = subprocess.run([krepp, "--num-threads", "4", "dist", "-i", d, "-q", "reads.fq"], cwd=out, check=True, capture_output=True, text=True)
        outs.append(sorted(r.stdout.splitlines()[2:]))
    assert outs[0] == outs[1] and len(outs[0]) > 3000
