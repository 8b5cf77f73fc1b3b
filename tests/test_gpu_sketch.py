"""GPU: `krepp sketch` -- the minimizer kernel with its HyperLogLog registers, the LSH position draw
and the writer.  The file krepp_b200 writes must be the file the UNMODIFIED reference writes, byte for byte, for every geometry,
with and without --seed, on single- and multi-sequence inputs with runs of N, short contigs and lower case; the estimates
against the oracle's restatement."""
import filecmp
import os
import subprocess

import pytest

import conftest
from conftest import needs_ref
from sketches import SKETCHES, SMALL, ref_seek
from test_seek_cpu import fasta_seqs

pytestmark = [pytest.mark.gpu, needs_ref]
EXE = os.path.join(conftest.ROOT, "krepp_b200", "_build", "krepp_b200")
REF = os.path.join(conftest.REF_DIR, "krepp")


def contigs_fasta(path):
    """A draft assembly: contigs with runs of N, one shorter than any window, one ending in a valid run shorter than the window
    right after an N (the reference's end-of-sequence emit), lower-case stretches, wrapped lines."""
    g = [fasta_seqs(os.path.join(SMALL, "genomes", f"G00000{i}.fna"))[0].decode() for i in (1, 2, 4)]
    parts = [g[0][:9000], g[0][9000:9400] + "N" * 30 + g[0][9430:15000].lower(), "ACGTACGTACGTACGTACGT", g[1][100:4000] + "NNNN" + g[1][4004:4033],
             g[2][:12000] + "N" + g[2][12001:12020], "N" * 50, g[1][6000:6040]]
    with open(path, "w") as f:
        for i, s in enumerate(parts):
            f.write(f">contig{i} some description\n")
            for j in range(0, len(s), 70):
                f.write(s[j:j + 70] + "\n")
    return [p.encode() for p in parts]


def both(args, inp, tmp_path, tag, pre=()):
    mine, ref = str(tmp_path / f"{tag}.mine.skc"), str(tmp_path / f"{tag}.ref.skc")
    subprocess.run([REF, *pre, "sketch", "-i", inp, "-o", ref, *args], check=True, capture_output=True)
    r = subprocess.run([EXE, *pre, "sketch", "-i", inp, "-o", mine, *args], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    assert "Total number of k-mers included in the sketch:" in r.stderr and "Subsampling rate (rho) is:" in r.stderr
    return mine, ref


@pytest.mark.parametrize("label,genome,args", SKETCHES, ids=[s[0] for s in SKETCHES])
def test_sketch_file_equals_the_reference(label, genome, args, tmp_path):
    mine, ref = both(args, os.path.join(SMALL, "genomes", genome + ".fna"), tmp_path, label)
    assert filecmp.cmp(mine, ref, shallow=False), label


def test_sketch_of_contigs_seeded_and_gzip(tmp_path):
    import gzip
    import oracle_lib as O
    import krepp_b200
    fa = str(tmp_path / "contigs.fa")
    parts = contigs_fasta(fa)
    for tag, args, pre in (("c_default", [], ()), ("c_seed", ["-k", "23", "-w", "30", "-h", "9", "-m", "5", "-r", "2"], ("--seed", "11")),
                           ("c_w_eq_k", ["-k", "24", "-w", "24", "-h", "8", "-m", "2", "-r", "0", "--no-frac"], ())):
        mine, ref = both(args, fa, tmp_path, tag, pre)
        assert filecmp.cmp(mine, ref, shallow=False), tag
        meta = O.read_sketch_file(mine)
        keys, rho = O.oracle_sketch_table(meta, parts)
        assert rho == meta["rho"] and len(keys) == len(meta["enc"])
        g = krepp_b200.Index.geometry(meta["k"], meta["w"], meta["h"], meta["m"], meta["r"], bool(meta["frac"]), seed=11 if tag == "c_seed" else None)
        n1, n2 = g.sequence_rho(parts)
        assert n2 / n1 == meta["rho"]
        assert (g.extract_mers(parts) == keys).all()
        g.close()
    gz = str(tmp_path / "contigs.fa.gz")
    with open(fa, "rb") as f, gzip.open(gz, "wb") as z:
        z.write(f.read())
    mine, ref = both([], gz, tmp_path, "c_gz")
    assert filecmp.cmp(mine, ref, shallow=False)
    # the sketch just written serves `seek`
    q = os.path.join(SMALL, "reads.fq")
    out = subprocess.run([EXE, "seek", "-i", mine, "-q", q], capture_output=True, text=True, check=True).stdout.splitlines()
    assert sorted(out) == sorted(ref_seek(ref, q))


def test_sketch_argument_errors(tmp_path):
    g = os.path.join(SMALL, "genomes", "G000000.fna")
    for args, msg in ((["-k", "26", "-w", "20", "-h", "10"], "The minimum minimizer window size (-w) is k (-k)."),
                      (["-k", "30", "-w", "36", "-h", "10"], "For compact k-mer encodings, h must be >= k-16."),
                      (["-k", "40"], "not in range [19 - 31]")):
        r = subprocess.run([EXE, "sketch", "-i", g, "-o", str(tmp_path / "x.skc"), *args], capture_output=True, text=True)
        q = subprocess.run([REF, "sketch", "-i", g, "-o", str(tmp_path / "y.skc"), *args], capture_output=True, text=True)
        assert r.returncode != 0 and msg in r.stderr, r.stderr
        assert q.returncode != 0 and msg in q.stderr + q.stdout, q.stderr
