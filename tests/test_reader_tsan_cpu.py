"""Host threads under ThreadSanitizer.  (1) The host reader's chunk-parallel FASTQ framing (krepp_reader_set_threads; row a1 / f1) under ThreadSanitizer: host_io.cpp and
index_image.cpp compiled with -fsanitize=thread into a small driver (tests/native/reader_tsan.cpp) that reads a FASTQ file with
one and with eight threads and compares the records.  (2) The index loader (threads pread the k-mer table and flatten the colour
lists level by level): the golden index whole and as three bucket-range shards.  (3) The library writer's table stage (threads
over ranges of k-mers, each writing the offsets of the rows that end in its range).  A data race makes the driver exit non-zero
(TSAN_OPTIONS=halt_on_error)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "reader_tsan.cpp")
CSRC = os.path.join(ROOT, "krepp_b200", "csrc")
OUT = os.path.join(ROOT, "oracle", "_build", "reader_tsan")


@pytest.fixture(scope="module")
def driver():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-I/usr/local/cuda/include", "-o", OUT, SRC,
                        os.path.join(CSRC, "host_io.cpp"), os.path.join(CSRC, "index_image.cpp"), os.path.join(CSRC, "library_writer.cpp"), "-lz", "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer build failed here: " + r.stderr[-300:])
    return OUT


def write_fastq(path, n, rng, ragged=False):
    acgt = np.frombuffer(b"ACGTN", np.uint8)
    with open(path, "wb") as f:
        for i in range(n):
            ln = int(rng.integers(1, 300)) if ragged else 150
            s = acgt[rng.integers(0, 5 if ragged else 4, ln)].tobytes()
            q = bytes(rng.integers(33, 74, ln).astype(np.uint8))  # quality lines may start with '@' or '+'
            f.write(b"@r%d some comment\n%s\n+\n%s\n" % (i, s, q))


@pytest.mark.parametrize("ragged", [False, True], ids=["150bp", "ragged"])
def test_parallel_reader_has_no_data_race(driver, ragged, tmp_path):
    rng = np.random.default_rng(2 + ragged)
    fq = str(tmp_path / "reads.fq")
    write_fastq(fq, 60000, rng, ragged)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1:exitcode=66")
    for batch in ("4096", "50000"):
        r = subprocess.run([driver, fq, "8", batch], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        assert "60000 records, 8 threads: same as one thread" in r.stdout


def test_index_loader_has_no_data_race(driver):
    idx = os.path.join(ROOT, "tests", "golden", "small", "index")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1:exitcode=66")
    for nshards in ("1", "3"):
        r = subprocess.run([driver, "--load", idx, nshards], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        assert f"{nshards} shard(s) hold 24528 entries" in r.stdout
    toy = os.path.join(ROOT, "oracle", "_ref", "toy", "index_toy")
    if os.path.isdir(toy):  # 6.9 M entries: the table is read by several threads
        r = subprocess.run([driver, "--load", toy], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0 and "1 shard(s) hold 6934548 entries" in r.stdout, (r.stdout + r.stderr)[-3000:]


def test_library_writer_has_no_data_race(driver, tmp_path):
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1:exitcode=66")
    r = subprocess.run([driver, "--write", str(tmp_path / "index"), str(6 << 20)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "6291456 k-mers written, 16 colour ids, offsets as counted sequentially" in r.stdout
