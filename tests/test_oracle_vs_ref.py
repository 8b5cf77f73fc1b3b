"""CPU tests that need the compiled reference (oracle/_ref, built by oracle/Makefile where /root/reference exists):
the oracle against LIVE outputs of the unmodified reference on the toy index (config 1) and on fresh synthetic reads."""
import os
import subprocess

import pytest

from conftest import REF_DIR, TOY_DIR, needs_ref

pytestmark = [needs_ref]


def _dump(fq, *extra):
    import oracle_lib as O
    txt = subprocess.run([os.path.join(REF_DIR, "ref_dump"), os.path.join(TOY_DIR, "index_toy"), fq, "--lookups", "--place", *extra],
                         capture_output=True, text=True, check=True).stdout
    return O.parse_ref_dump(txt)


def _compare(dump, reads, **pk):
    import oracle_lib as O
    ix = O.OracleIndex(os.path.join(TOY_DIR, "index_toy"))
    p = O.default_params(want_lookups=1, want_place=1, no_filter=0, **pk)
    n = 0
    for i, s in enumerate(reads):
        o, r = ix.query(s, p), dump["reads"][i]
        for key in ("onmers", "wn", "hdist_filt", "lookups"):
            assert o[key] == r[key], (i, key)
        key_m = lambda m: (m["strand"], m["leaf_se"], m["match"], m["hdist_min"], m["rho"], m["hist"])
        assert [key_m(m) for m in o["minfo"]] == [key_m(m) for m in r["minfo"]], i
        key_s = lambda s_: (s_["leaf_se"], s_["strand"], s_["d"], s_["v"], s_["chisq"], s_["is_closest"])
        assert [key_s(x) for x in o["sel"]] == [key_s(x) for x in r["sel"]], i
        assert [tuple(q.values()) for q in o["place"]] == [tuple(q.values()) for q in r["place"]], i
        n += len(r["place"])
    return n


def test_toy_query_config1():
    with open(os.path.join(TOY_DIR, "query_toy.fq"), "rb") as f:
        reads = f.read().split(b"\n")[1::4]
    d = _dump(os.path.join(TOY_DIR, "query_toy.fq"))
    assert d["info"]["mask_hash_bp"] == 0x00333C0C3003CCF0 and d["info"]["mask_drop_lr"] == 0x029DBE53029DBE53
    assert _compare(d, reads) in (386, 387)  # the reference itself emits 386 or 387 rows


def test_synthetic_reads_and_other_thresholds(tmp_path):
    import synth
    seq, offs = synth.load_packed(os.path.join(TOY_DIR, "genomes.npz"))
    m = synth.sample_reads(seq, offs, 3000, seed=4)
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, m)
    reads = [r.tobytes() for r in m]
    _compare(_dump(fq), reads)
    _compare(_dump(fq, "--hdist-th", "6", "--tau", "3"), reads, hdist_th=6, tau=3)


def test_dist_cli_config1():
    ref = subprocess.run([os.path.join(REF_DIR, "krepp"), "dist", "-i", os.path.join(TOY_DIR, "index_toy"), "-q", os.path.join(TOY_DIR, "query_toy.fq")],
                         capture_output=True, text=True, check=True).stdout.splitlines()[2:]
    ora = subprocess.run([os.path.join(os.path.dirname(REF_DIR), "_build", "krepp_oracle"), "dist", os.path.join(TOY_DIR, "index_toy"),
                          os.path.join(TOY_DIR, "query_toy.fq")], capture_output=True, text=True, check=True).stdout.splitlines()[1:]
    assert len(ref) == 413 and sum(1 for l in ref if l.endswith("NaN")) == 5  # SURVEY.md 8c known answers
    assert sorted(ref) == sorted(ora)
