"""CPU: the command line's argument handling (everything that is decided before a GPU is touched) against the reference's own
CLI11 front end (ref src/krepp.cpp:593-716): same required options, same ranges, same validation messages, non-zero exit."""
import os
import subprocess

import pytest

from conftest import GOLDEN_DIR, REF_DIR, ROOT, needs_ref

CLI = os.path.join(ROOT, "krepp_b200", "_build", "krepp_b200")
S = os.path.join(GOLDEN_DIR, "small")
IDX, FQ = os.path.join(S, "index"), os.path.join(S, "reads.fq")

CASES = [
    ([], "A subcommand is required"),
    (["dist"], "--query is required"),
    (["dist", "-q", FQ], "--index-dir is required"),
    (["dist", "-q", "nope.fq", "-i", IDX], "File does not exist: nope.fq"),
    (["dist", "-q", FQ, "-i", "nope_dir"], "Directory does not exist: nope_dir"),
    (["place", "-i", IDX, "-q", FQ, "--tau", "9"], "The threshold tau must be less than HD threshold --hdist-th!"),
    (["dist", "-i", IDX, "-q", FQ, "--dist-max", "0.9"], "not in range [1e-08 - 0.33]"),
    (["dist", "-i", IDX, "-q", FQ, "--bogus"], "The following argument was not expected: --bogus"),
    (["place", "-i", IDX, "-q", FQ, "-t", os.path.join(S, "tree.nwk"), "-l", os.path.join(S, "input_map.tsv")], "--nwk-file excludes --lineage-file"),
]


@pytest.fixture(scope="module", autouse=True)
def built():
    import krepp_b200
    krepp_b200.build_library()


@pytest.mark.parametrize("args,msg", CASES, ids=[" ".join(a[:2]) + f" #{i}" for i, (a, _) in enumerate(CASES)])
def test_argument_errors(args, msg):
    r = subprocess.run([CLI, *args], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and msg in r.stderr, r.stderr
    if os.path.exists(os.path.join(REF_DIR, "krepp")):  # the reference says the same thing
        q = subprocess.run([os.path.join(REF_DIR, "krepp"), *args], capture_output=True, text=True, timeout=60)
        assert q.returncode != 0 and msg in (q.stderr + q.stdout), q.stderr


def test_subcommands_outside_the_query_path_are_refused_by_name():
    r = subprocess.run([CLI, "inspect", "-i", "x"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "Subcommand 'inspect' is not part of the GPU query path" in r.stderr
    r = subprocess.run([CLI, "index", "-i", FQ], capture_output=True, text=True, timeout=60)             # `index` is built (row f3)
    assert r.returncode != 0 and "--index-dir is required" in r.stderr
    r = subprocess.run([CLI, "index", "-o", "/tmp/x_idx", "-i", os.path.join(S, "no_such_map.tsv")], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "--input-file: File does not exist" in r.stderr
    r = subprocess.run([CLI, "index", "-o", "/tmp/x_idx", "-i", FQ, "-t", os.path.join(S, "no_such_tree.nwk")], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "--nwk-file: File does not exist" in r.stderr
    r = subprocess.run([CLI, "index", "-o", "/tmp/x_idx", "-i", FQ, "--sdust-t", "20", "--sdust-w", "64"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "dustmasker" in r.stderr
    r = subprocess.run([CLI, "place", "-i", IDX, "-q", FQ, "-l", os.path.join(S, "no_such_lineages.tsv")], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "--lineage-file: File does not exist" in r.stderr
    r = subprocess.run([CLI, "place", "-i", IDX, "-q", FQ, "-t", os.path.join(S, "no_such_tree.nwk")], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "--nwk-file: File does not exist" in r.stderr
    r = subprocess.run([CLI, "seek", "-i", IDX, "-q", FQ], capture_output=True, text=True, timeout=60)   # seek takes a sketch FILE
    assert r.returncode != 0 and "--sketch-path: File does not exist" in r.stderr
    r = subprocess.run([CLI, "seek", "-q", FQ], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "--sketch-path is required" in r.stderr
    r = subprocess.run([CLI, "sketch", "-i", FQ], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "--output-path is required" in r.stderr
    r = subprocess.run([CLI, "sketch", "-i", FQ, "-o", "/tmp/x.skc", "-k", "26", "-w", "20", "-h", "10"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "The minimum minimizer window size (-w) is k (-k)." in r.stderr
    r = subprocess.run([CLI, "--help"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "--shard-index" in r.stdout and "--num-gpus" in r.stdout


def test_no_gpu_means_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([CLI, "dist", "-i", IDX, "-q", FQ], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr and r.stdout.strip() == ""
