#!/usr/bin/env python
"""Regenerates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/krepp + oracle/_ref/ref_dump).

Run in the container that has /root/reference (after `make -C oracle ref`):   python tests/golden/make_golden.py
Produces under tests/golden/small/:
  genomes/*.fna, input_map.tsv, tree.nwk   8 synthetic genomes (24 kbp, JC69 down a random binary tree, seed 11)
  index/                                   `krepp index -k 21 -w 25 -h 7` built by the reference (default seed)
  reads.fq                                 240 reads: 150 bp sampled/mutated (seed 5) + hand-made edge cases
  ref_dump.txt.gz                          stage dump (lookups, histograms, solves, deterministic summarize/place)
  ref_dist.tsv, ref_place.jplace           raw CLI outputs of the reference (row order as the reference emitted it)
"""
import gzip
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tools")]
import numpy as np  # noqa: E402
import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden", "small")


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(os.path.join(OUT, "genomes"))
    names, seqs, nwk = synth.random_genomes(8, 24000, seed=11, depth_blen=0.03)
    with open(os.path.join(OUT, "tree.nwk"), "w") as f:
        f.write(nwk + "\n")
    with open(os.path.join(OUT, "input_map.tsv"), "w") as f:
        for nm, s in zip(names, seqs):
            s = s.copy()
            if nm == names[3]:
                s[5000:5040] = ord("N")  # an N run inside a genome (ring-buffer quirk of extract_mers)
            synth.write_fasta(os.path.join(OUT, "genomes", nm + ".fna"), nm + "_c1", s)
            f.write(f"{nm}\t./genomes/{nm}.fna\n")
    subprocess.run([os.path.join(REF, "krepp"), "index", "-k", "21", "-w", "25", "-h", "7", "-o", "index", "-i", "input_map.tsv", "-t", "tree.nwk"],
                   cwd=OUT, check=True, stderr=subprocess.DEVNULL)
    os.remove(os.path.join(OUT, "index", "metadata-m4r1-frac.txt"))  # carries a date
    # reads
    allseq = np.concatenate(seqs)
    offs = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    reads = [r.tobytes() for r in synth.sample_reads(allseq, offs, 200, read_len=150, max_sub=0.12, seed=5)]
    long_reads = [r.tobytes() for r in synth.sample_reads(allseq, offs, 4, read_len=700, max_sub=0.03, seed=6)]
    r0 = reads[0]
    edge = [b"A", r0[:20], r0[:21], r0[:22], b"N" * 80, r0[:70].lower() + r0[70:], r0[:50] + b"N" + r0[51:], r0[:30] + b"NNNN" + r0[34:100],
            b"ACGT" * 30, b"A" * 120, r0[:149] + b"*", r0.replace(b"A", b"R")]
    edge += [long_reads[0], long_reads[1][:257], long_reads[2][:128], long_reads[3][:129]]
    rng = np.random.default_rng(9)
    for _ in range(20):
        ln = int(rng.integers(15, 320))
        src = long_reads[int(rng.integers(0, 4))][:ln]
        edge.append(src)
    synth.write_fastq(os.path.join(OUT, "reads.fq"), reads + edge)
    dump = subprocess.run([os.path.join(REF, "ref_dump"), "index", "reads.fq", "--lookups", "--place"], cwd=OUT, check=True,
                          capture_output=True, text=True).stdout
    with gzip.open(os.path.join(OUT, "ref_dump.txt.gz"), "wt", compresslevel=9) as f:
        f.write(dump)
    dist = subprocess.run([os.path.join(REF, "krepp"), "dist", "-i", "index", "-q", "reads.fq"], cwd=OUT, check=True, capture_output=True, text=True).stdout
    with open(os.path.join(OUT, "ref_dist.tsv"), "w") as f:
        f.write("".join(l + "\n" for l in dist.splitlines()[2:]))
    place = subprocess.run([os.path.join(REF, "krepp"), "place", "-i", "index", "-q", "reads.fq"], cwd=OUT, check=True, capture_output=True, text=True).stdout
    with open(os.path.join(OUT, "ref_place.jplace"), "w") as f:
        f.write(place)
    for root, _, files in os.walk(OUT):
        for fn in files:
            print(f"{os.path.getsize(os.path.join(root, fn)):9d}  {os.path.relpath(os.path.join(root, fn), OUT)}")


if __name__ == "__main__":
    main()
