"""Sketches of single golden genomes built on the fly by the UNMODIFIED reference (`oracle/_ref/krepp sketch`) for `krepp seek`
(SURVEY.md 8 row f4): the reference's default geometry (k 26, w 32, h 10, m 4, r 1, frac) and others."""
import os
import subprocess

from conftest import GOLDEN_DIR, REF_DIR

SMALL = os.path.join(GOLDEN_DIR, "small")
# (label, genome, krepp sketch arguments)
SKETCHES = [
    ("default", "G000000", []),
    ("k29_w35_h13_m8r2_nofrac", "G000003", ["-k", "29", "-w", "35", "-h", "13", "-m", "8", "-r", "2", "--no-frac"]),
    ("k21_w21_h7_m3r1", "G000005", ["-k", "21", "-w", "21", "-h", "7", "-m", "3", "-r", "1"]),
    ("k28_w33_h12_m4r3_nofrac", "G000006", ["-k", "28", "-w", "33", "-h", "12", "-m", "4", "-r", "3", "--no-frac"]),
]


def build_sketch(label, genome, args, tmp_root):
    path = os.path.join(str(tmp_root), "sketch_" + label + ".skc")
    if not os.path.exists(path):
        subprocess.run([os.path.join(REF_DIR, "krepp"), "sketch", "-i", os.path.join(SMALL, "genomes", genome + ".fna"), "-o", path, *args], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return path


def ref_seek(path, query, th=4):
    out = subprocess.run([os.path.join(REF_DIR, "krepp"), "seek", "-i", path, "-q", query, "--hdist-th", str(th)], capture_output=True, text=True, check=True).stdout
    return out.splitlines()
