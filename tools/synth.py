"""Synthetic inputs for tests and bench.

Pure numpy; nothing here is on the product path.  Genomes are kept as uint8 ASCII arrays.
"""
from __future__ import annotations

import glob
import gzip
import os

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.full(256, ord("N"), dtype=np.uint8)
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[a] = b


def read_fasta(path: str) -> list[tuple[str, np.ndarray]]:
    """Returns [(name, uint8 ASCII sequence)] for every record of a (possibly gzipped) FASTA file."""
    op = gzip.open if path.endswith(".gz") else open
    recs, name, chunks = [], None, []
    with op(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    recs.append((name, np.frombuffer(b"".join(chunks), dtype=np.uint8)))
                name, chunks = line[1:].split()[0].decode(), []
            else:
                chunks.append(line.strip())
    if name is not None:
        recs.append((name, np.frombuffer(b"".join(chunks), dtype=np.uint8)))
    return recs


def pack_genomes(fasta_dir: str, out_npz: str) -> None:
    """Concatenates all contigs of all *.fna files under fasta_dir into one uint8 array + contig offsets."""
    seqs, offs, names = [], [0], []
    for f in sorted(glob.glob(os.path.join(fasta_dir, "*.fna"))):
        for name, s in read_fasta(f):
            seqs.append(s)
            offs.append(offs[-1] + len(s))
            names.append(os.path.basename(f)[:-4] + "|" + name)
    np.savez_compressed(out_npz, seq=np.concatenate(seqs), offsets=np.asarray(offs, dtype=np.int64))


def load_packed(npz: str) -> tuple[np.ndarray, np.ndarray]:
    z = np.load(npz)
    return z["seq"], z["offsets"]


def sample_reads(seq: np.ndarray, offsets: np.ndarray, n: int, read_len: int = 150, max_sub: float = 0.15,
                 seed: int = 1) -> np.ndarray:
    """Config-2 style reads: start uniform over all valid positions (=> contig chosen with probability proportional to
    length), per-read substitution rate U(0, max_sub) applied i.i.d. per base with a replacement drawn uniformly from
    ACGT, strand flipped with probability 0.5.  Returns an (n, read_len) uint8 ASCII matrix (upper-cased)."""
    rng = np.random.default_rng(seed)
    lens = np.diff(offsets)
    valid = np.maximum(lens - read_len + 1, 0)
    cum = np.concatenate([[0], np.cumsum(valid)])
    pick = rng.integers(0, cum[-1], size=n)
    contig = np.searchsorted(cum, pick, side="right") - 1
    start = offsets[contig] + (pick - cum[contig])
    reads = seq[start[:, None] + np.arange(read_len)[None, :]].copy()
    reads &= 0xDF  # upper-case ASCII letters
    d = rng.uniform(0.0, max_sub, size=n)
    mask = rng.random((n, read_len)) < d[:, None]
    reads[mask] = _ACGT[rng.integers(0, 4, size=int(mask.sum()))]
    flip = rng.random(n) < 0.5
    reads[flip] = _COMP[reads[flip][:, ::-1]]
    return reads


def write_fastq(path: str, reads, names=None) -> None:
    """reads: (n, L) uint8 matrix or list of bytes."""
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            s = r.tobytes() if isinstance(r, np.ndarray) else r
            nm = names[i] if names is not None else f"r{i}"
            f.write(b"@" + nm.encode() + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")


def random_genomes(n_genomes: int, length: int, seed: int = 7, depth_blen: float = 0.02):
    """Config-3 style genomes: root i.i.d. uniform ACGT, evolved down a random rooted binary tree with JC69
    substitutions only, branch lengths Exp(mean depth_blen).  Returns (names, [uint8 ASCII], newick)."""
    rng = np.random.default_rng(seed)

    def jc_mutate(s, t):
        p = 0.75 * (1.0 - np.exp(-4.0 * t / 3.0))  # probability that a site differs after time t
        m = rng.random(len(s)) < p
        out = s.copy()
        k = int(m.sum())
        if k:
            out[m] = (out[m] + rng.integers(1, 4, size=k).astype(np.uint8)) & 3
        return out

    names, seqs = [], []
    counter = [0]

    def grow(s, n_leaves):
        if n_leaves == 1:
            nm = f"G{counter[0]:06d}"
            counter[0] += 1
            names.append(nm)
            seqs.append(_ACGT[s])
            return nm
        left = int(rng.integers(1, n_leaves))
        parts = []
        for nl in (left, n_leaves - left):
            t = float(rng.exponential(depth_blen)) + 1e-4
            parts.append(f"{grow(jc_mutate(s, t), nl)}:{t:.6f}")
        return "(" + ",".join(parts) + ")"

    root = rng.integers(0, 4, size=length).astype(np.uint8)
    # the reference's Newick reader indexes past its token vector for a root without label AND length
    # (src/phytree.cpp:171-187), so the root always gets both
    nwk = grow(root, n_genomes) + "root:0.0;"
    return names, seqs, nwk


def write_fasta(path: str, name: str, seq: np.ndarray, width: int = 80) -> None:
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        b = seq.tobytes()
        for i in range(0, len(b), width):
            f.write(b[i:i + width] + b"\n")
