"""GPU: `krepp index` -- leaf tables and rho per genome on the GPU (minimizer kernel), the union of all leaf
tables as one sort, the distinct reference sets, then the host's colour record and writers.  The library `krepp_b200 index` writes
must be the library the UNMODIFIED reference writes from the same inputs up to the numbering of colours above the tree nodes
(tests/libraries.py), the reference's `krepp dist` must answer alike from either, and so must `krepp_b200 dist`."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

import conftest
import krepp_b200
from conftest import needs_ref
from libraries import assert_same_library, colour_leaves, read_library
from test_gpu_sketch import contigs_fasta
from test_seek_cpu import fasta_seqs
from variants import SMALL

pytestmark = [pytest.mark.gpu, needs_ref]
EXE = os.path.join(conftest.ROOT, "krepp_b200", "_build", "krepp_b200")
REF = os.path.join(conftest.REF_DIR, "krepp")
READS = os.path.join(SMALL, "reads.fq")


def workdir(tmp_path, tag):
    w = str(tmp_path / tag)
    os.makedirs(w)
    shutil.copytree(os.path.join(SMALL, "genomes"), os.path.join(w, "genomes"))
    shutil.copy(os.path.join(SMALL, "input_map.tsv"), w)
    shutil.copy(os.path.join(SMALL, "tree.nwk"), w)
    return w


def build_both(w, args, with_tree=True, pre=(), threads="3"):
    tree = ["-t", "tree.nwk"] if with_tree else []
    subprocess.run([REF, *pre, "index", *args, "-o", "ref_index", "-i", "input_map.tsv", *tree], cwd=w, check=True, capture_output=True)
    r = subprocess.run([EXE, "--num-threads", threads, "--verbose", *pre, "index", *args, "-o", "my_index", "-i", "input_map.tsv", *tree], cwd=w, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Finished indexing, elapsed:" in r.stderr and "Done converting & saving, elapsed:" in r.stderr
    return os.path.join(w, "my_index"), os.path.join(w, "ref_index")


def dist_lines(exe, index_dir, extra=()):
    out = subprocess.run([exe, "dist", "-i", index_dir, "-q", READS, *extra], capture_output=True, text=True, check=True).stdout
    return sorted(l for l in out.splitlines() if not l.startswith("#"))


def check_pair(mine_dir, ref_dir):
    mine, ref = read_library(mine_dir), read_library(ref_dir)
    assert_same_library(mine, ref)
    want = dist_lines(REF, ref_dir)
    assert dist_lines(REF, mine_dir) == want       # the reference reads this library and answers as from its own
    assert dist_lines(EXE, mine_dir) == want       # and the GPU query path on the GPU-built library
    return mine, ref


@pytest.mark.parametrize("label,args,pre", [
    ("k21_h7", ["-k", "21", "-w", "25", "-h", "7"], ()),
    ("k25_h9_m5r3_seed", ["-k", "25", "-w", "31", "-h", "9", "-m", "5", "-r", "3"], ("--seed", "7")),
    ("k21_h7_m3r2_nofrac", ["-k", "21", "-w", "25", "-h", "7", "-m", "3", "-r", "2", "--no-frac"], ()),
], ids=lambda v: v if isinstance(v, str) else "")
def test_library_equals_the_reference(label, args, pre, tmp_path):
    w = workdir(tmp_path, label)
    mine, ref = check_pair(*build_both(w, args, pre=pre))
    assert mine["nsubsets"] <= ref["nsubsets"]


def test_library_without_a_guide_tree(tmp_path):
    w = workdir(tmp_path, "treeless")
    mine_dir, ref_dir = build_both(w, ["-k", "21", "-w", "25", "-h", "7"], with_tree=False)
    mine, ref = check_pair(mine_dir, ref_dir)
    assert mine["tree"] is None


def test_library_of_draft_assemblies(tmp_path):
    """Contigs with runs of N, contigs shorter than a window, the end-of-sequence emit, lower case, gzip; a leaf of the tree without
    a genome and a reference id the tree does not have."""
    w = workdir(tmp_path, "drafts")
    contigs_fasta(os.path.join(w, "genomes", "G000001.fna"))
    with open(os.path.join(w, "genomes", "G000002.fna"), "rb") as f, gzip.open(os.path.join(w, "genomes", "G000002.fna.gz"), "wb") as z:
        z.write(f.read())
    rows = [l.rstrip("\n").split("\t") for l in open(os.path.join(w, "input_map.tsv"))]
    with open(os.path.join(w, "input_map.tsv"), "w") as f:
        for name, path in rows:
            if name == "G000005":
                f.write("GXXXXXX\t" + path + "\n")  # G000005 stays on the tree without a genome; GXXXXXX is on no tree
            elif name == "G000002":
                f.write(name + "\t" + path + ".gz\n")
            else:
                f.write(name + "\t" + path + "\n")
    mine, ref = check_pair(*build_both(w, ["-k", "21", "-w", "25", "-h", "7"]))
    assert b"GXXXXXX\n" in mine["reflist"]


def test_builder_through_the_c_abi(tmp_path):
    """krepp_builder_add_genome / _union / _write called directly: counts against a numpy restatement of the union over the
    GPU's own leaf tables, rho against krepp_sequence_rho, and the written library against the reference's."""
    w = workdir(tmp_path, "capi")
    subprocess.run([REF, "index", "-k", "21", "-w", "25", "-h", "7", "-o", "ref_index", "-i", "input_map.tsv", "-t", "tree.nwk"], cwd=w, check=True, capture_output=True)
    ref = read_library(os.path.join(w, "ref_index"))
    names = [l.split("\t")[0] for l in open(os.path.join(w, "input_map.tsv"))]
    g = krepp_b200.Index.geometry(21, 25, 7, 4, 1, True)
    b = krepp_b200.LibraryBuilder(g, open(os.path.join(w, "tree.nwk")).read(), names)
    pairs = set()
    for nm in reversed(names):  # any order
        seqs = fasta_seqs(os.path.join(w, "genomes", nm + ".fna"))
        n, rho = b.add_genome(nm, seqs)
        keys = g.extract_mers(seqs)
        n1, n2 = g.sequence_rho(seqs)
        assert n == len(keys) and rho == n2 / n1
        pairs.update((int(k), b.leaf_rank(nm)) for k in keys)
    with pytest.raises(krepp_b200.KreppError, match="added before"):
        b.add_genome(names[0], [b"ACGT" * 30])
    nk, nsets = b.union()
    by_key = {}
    for k, r in pairs:
        by_key.setdefault(k, []).append(r)
    assert nk == len(by_key) == ref["nkmers"]
    assert nsets == len({tuple(sorted(v)) for v in by_key.values()})
    out = os.path.join(w, "my_index")
    assert b.write(out)[0] == nk
    assert_same_library(read_library(out), ref)
    b.close(); g.close()


def test_index_argument_errors(tmp_path):
    w = workdir(tmp_path, "errors")
    run = lambda *a: subprocess.run([EXE, *a], cwd=w, capture_output=True, text=True)
    r = run("index", "-o", "x", "-i", "input_map.tsv", "-t", "tree.nwk", "-k", "21", "-w", "20", "-h", "7")
    assert r.returncode == 1 and "The minimum minimizer window size (-w) is k (-k)." in r.stderr
    with open(os.path.join(w, "bad_map.tsv"), "w") as f:
        f.write("G000000\n")
    r = run("index", "-o", "x", "-i", "bad_map.tsv", "-t", "tree.nwk")
    assert r.returncode == 1 and "Failed to read the reference name to path/URL mapping!" in r.stderr
    with open(os.path.join(w, "gone_map.tsv"), "w") as f:
        f.write("G000000\t./genomes/none.fna\n")
    r = run("index", "-o", "x", "-i", "gone_map.tsv", "-t", "tree.nwk", "-k", "21", "-w", "25", "-h", "7")
    assert r.returncode == 1 and "Failed to open the file at ./genomes/none.fna" in r.stderr
