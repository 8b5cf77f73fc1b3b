"""Index geometries other than the default -k 27 -w 35 -h 11 -m 4 -r 1 --frac, built on the fly by the UNMODIFIED reference
(`oracle/_ref/krepp index`) from the committed golden genomes: SURVEY.md 8 rows a4/a5 (hash geometry, row addressing for
frac / no-frac and m that is not a power of two) and row a16 (metadata of every suffix form).  Shared by the CPU test
(oracle against the reference's own outputs) and the GPU test (CUDA path against the oracle)."""
import os
import shutil
import subprocess

from conftest import GOLDEN_DIR, REF_DIR

SMALL = os.path.join(GOLDEN_DIR, "small")
# (label, krepp index arguments)
VARIANTS = [
    ("k19_h5_m2r0_frac", ["-k", "19", "-w", "23", "-h", "5", "-m", "2", "-r", "0"]),
    ("k21_h7_m3r2_nofrac", ["-k", "21", "-w", "25", "-h", "7", "-m", "3", "-r", "2", "--no-frac"]),
    ("k25_h9_m5r3_frac", ["-k", "25", "-w", "31", "-h", "9", "-m", "5", "-r", "3"]),
    # h = k - 16 is forced for k > 27 (32-bit residual encoding), so the row space is 4^15 / 4^13 wide: a large modulo keeps the
    # offset array at tens of MB (the reference's own default, k29 h13 m4 r1, has 2^25 rows)
    ("k31_h15_m32r7_nofrac", ["-k", "31", "-w", "37", "-h", "15", "-m", "32", "-r", "7", "--no-frac"]),
    ("k29_h13_m16r1_frac", ["-k", "29", "-w", "35", "-h", "13", "-m", "16", "-r", "1"]),
    ("k20_h6_m1r0", ["-k", "20", "-w", "24", "-h", "6", "-m", "1", "-r", "0"]),
]


# An index built WITHOUT a backbone tree: the reference writes reflist-* instead of tree-* and generates a balanced tree over
# the names at load time (ref src/index.cpp:3-27, src/phytree.cpp:217-253); `dist` works on it, `place` needs a tree.
TREELESS = ("treeless_k21_h7", ["-k", "21", "-w", "25", "-h", "7"])


def build(label, args, tmp_root, with_tree=True):
    """Runs the reference's `krepp index` in a scratch copy of the golden genomes; returns the index directory."""
    work = os.path.join(str(tmp_root), label)
    if not os.path.isdir(os.path.join(work, "index")):
        os.makedirs(work, exist_ok=True)
        for item in ("genomes", "input_map.tsv", "tree.nwk"):
            src, dst = os.path.join(SMALL, item), os.path.join(work, item)
            (shutil.copytree if os.path.isdir(src) else shutil.copy)(src, dst)
        subprocess.run([os.path.join(REF_DIR, "krepp"), "index", *args, "-o", "index", "-i", "input_map.tsv", *(["-t", "tree.nwk"] if with_tree else [])],
                       cwd=work, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.join(work, "index")


def build_partials(tmp_root, rs=("0", "2", "3"), frac=False):
    """A directory holding several partial libraries: separate `krepp index` runs (m = 4, one residue each, or with --frac the
    residues 0..r) into ONE -o directory; every partial has its own table, colour record and rho (ref src/krepp.cpp:66-108)."""
    work = os.path.join(str(tmp_root), "partials_" + "_".join(rs) + ("_frac" if frac else ""))
    if not os.path.isdir(os.path.join(work, "index")):
        os.makedirs(work, exist_ok=True)
        for item in ("genomes", "input_map.tsv", "tree.nwk"):
            src, dst = os.path.join(SMALL, item), os.path.join(work, item)
            if not os.path.exists(dst):
                (shutil.copytree if os.path.isdir(src) else shutil.copy)(src, dst)
        for r in rs:
            subprocess.run([os.path.join(REF_DIR, "krepp"), "index", "-k", "21", "-w", "25", "-h", "7", "-m", "4", "-r", r, *([] if frac else ["--no-frac"]), "-o", "index",
                            "-i", "input_map.tsv", "-t", "tree.nwk"], cwd=work, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.join(work, "index")
