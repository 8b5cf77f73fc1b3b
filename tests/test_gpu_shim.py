"""The binding of INTEGRATION.md section 2 compiled for real: oracle/_ref/krepp_gpu is the reference's own program (its CLI11
front end, QSeq reader, OpenMP batch tasks, output framing) with the two IBatch construction sites of src/krepp.cpp redirected to
GpuBatch (oracle/shim/gpubatch.hpp), a class with IBatch's interface over the C ABI.  Its output must be the stock reference
binary's: `dist` line for line, `place` read by read (where the closest reference is not tied, SURVEY.md section 0 fact 6, and
in any case identical to the krepp_b200 executable's, which applies the same fixed tie rule through the same library)."""
import json
import os
import subprocess

import pytest

import conftest

pytestmark = [pytest.mark.gpu]
SHIM = os.path.join(conftest.REF_DIR, "krepp_gpu")
REF = os.path.join(conftest.REF_DIR, "krepp")
EXE = os.path.join(conftest.ROOT, "krepp_b200", "_build", "krepp_b200")
SMALL = os.path.join(conftest.GOLDEN_DIR, "small")
needs_shim = pytest.mark.skipif(not os.path.exists(SHIM), reason="oracle/_ref/krepp_gpu not built (make -C oracle shim, needs /root/reference)")


def run(exe, *args):
    r = subprocess.run([exe, *args], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def rows(jplace_text):
    return {p["n"][0]: sorted(map(tuple, p["p"])) for p in json.loads(jplace_text)["placements"]}


@needs_shim
def test_shim_dist_equals_the_reference_output():
    idx, q = os.path.join(SMALL, "index"), os.path.join(SMALL, "reads.fq")
    got = run(SHIM, "--num-threads", "4", "dist", "-i", idx, "-q", q).splitlines()
    with open(os.path.join(SMALL, "ref_dist.tsv")) as f:
        want = f.read().splitlines()
    assert got[1] == "SEQ_ID\tREFERENCE_NAME\tDIST" and sorted(got[2:]) == sorted(want) and len(want) > 500
    # other modes, against the stock binary run here (reads with a tied closest reference may differ under --filter / --no-multi)
    if os.path.exists(REF):
        for extra in (["--dist-max", "0.05"], ["--hdist-th", "3"]):
            a = sorted(run(SHIM, "dist", "-i", idx, "-q", q, *extra).splitlines()[2:])
            b = sorted(run(REF, "dist", "-i", idx, "-q", q, *extra).splitlines()[2:])
            assert a == b, extra


@needs_shim
def test_shim_place_equals_the_executable_and_the_reference_where_untied():
    idx, q = os.path.join(SMALL, "index"), os.path.join(SMALL, "reads.fq")
    shim = rows(run(SHIM, "--num-threads", "3", "place", "-i", idx, "-q", q))
    mine = rows(run(EXE, "place", "-i", idx, "-q", q))
    assert shim == mine and len(shim) > 100
    with open(os.path.join(SMALL, "ref_place.jplace")) as f:
        ref = rows(f.read())
    same = sum(1 for k in ref if shim.get(k) == ref[k])
    assert set(shim) == set(ref) or abs(len(shim) - len(ref)) <= 10
    assert same >= 0.85 * len(ref), (same, len(ref))   # the rest are reads whose closest reference is tied in the reference's hash order
    tab = run(SHIM, "place", "-i", idx, "-q", q, "--tabular").splitlines()
    assert len(tab) == 408 + 3


@needs_shim
def test_shim_summarize_tables_agree():
    idx, q = os.path.join(SMALL, "index"), os.path.join(SMALL, "reads.fq")
    a = run(SHIM, "dist", "-i", idx, "-q", q, "--summarize").splitlines()[2:]
    b = run(EXE, "dist", "-i", idx, "-q", q, "--summarize").splitlines()[2:]
    key = lambda l: l.split("\t")[0]
    assert sorted(a, key=key) == sorted(b, key=key) and len(a) > 3


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/krepp not built")
def test_cli_modes_equal_the_reference_cli_on_untied_reads(tmp_path):
    """--no-multi, --filter, --dist-max, --summarize (dist) and --no-multi / --tabular (place) against the stock reference binary.
    These modes consult the closest reference, which the reference picks in hash-map order when several leaves tie (SURVEY.md
    section 0 fact 6), so the comparison runs on the reads whose closest reference is unique -- there every line must agree."""
    import oracle_lib as O
    idx, q = os.path.join(SMALL, "index"), os.path.join(SMALL, "reads.fq")
    o = O.OracleIndex(idx)
    sub = tmp_path / "untied.fq"
    kept = 0
    with open(q) as f, open(sub, "w") as g:
        while True:
            rec = [f.readline() for _ in range(4)]
            if not rec[0]:
                break
            sel = o.query(rec[1].strip().encode(), O.default_params(no_filter=0))["sel"]
            if sel:
                dmin = min(x["d"] for x in sel)
                if sum(1 for x in sel if x["d"] == dmin) != 1:
                    continue
            g.writelines(rec)
            kept += 1
    assert kept > 100
    for args in (["dist", "--no-multi"], ["dist", "--filter"], ["dist", "--dist-max", "0.05"], ["dist", "--filter", "--dist-max", "0.08", "--no-multi"],
                 ["dist", "--hdist-th", "2"], ["place", "--tabular"], ["place", "--tabular", "--no-multi"], ["place", "--tabular", "--no-filter"]):
        a = sorted(run(EXE, *args, "-i", idx, "-q", str(sub)).splitlines()[2:])
        b = sorted(run(REF, *args, "-i", idx, "-q", str(sub)).splitlines()[2:])
        assert a == b and len(a) >= 50, (args, len(a), len(b), [x for x in a if x not in set(b)][:3], [x for x in b if x not in set(a)][:3])
    for args in (["dist", "--summarize"], ["dist", "--summarize", "--dist-max", "0.1"], ["place", "--summarize"]):
        def table(exe):
            out = {}
            for l in run(exe, *args, "-i", idx, "-q", str(sub)).splitlines():
                t = l.split("\t")
                if l and not l.startswith("#") and len(t) >= 3 and t[-1][:1].isdigit():
                    out[tuple(t[:-2])] = [float(x) for x in t[-2:]]
            return out
        a, b = table(EXE), table(REF)
        assert a.keys() == b.keys() and len(a) > 3, (args, sorted(set(a) ^ set(b))[:5])
        for k in a:
            assert all(abs(x - y) <= 2e-5 for x, y in zip(a[k], b[k])), (args, k, a[k], b[k])


ALT_TREE = ("((G000000:0.01,G000002:0.02):0.03,((G000003:0.01,G000001:0.02):0.01,(G000004:0.05,(G000005:0.02,GXXXXXX:0.01):0.02):0.03):0.02,"
            "G000006:0.08)root;")


def untied_subset(tmp_path, kept):
    """The golden reads whose closest reference, among the references `kept`, is unique (ties are broken by container order in
    the reference), as a FASTQ file."""
    import oracle_lib as O
    idx, q = os.path.join(SMALL, "index"), os.path.join(SMALL, "reads.fq")
    o = O.OracleIndex(idx)
    sub = tmp_path / "untied.fq"
    n = 0
    with open(q) as f, open(sub, "w") as g:
        while True:
            rec = [f.readline() for _ in range(4)]
            if not rec[0]:
                break
            sel = [x for x in o.query(rec[1].strip().encode(), O.default_params(no_filter=0))["sel"] if o.name(x["leaf_se"]) in kept]
            if sel:
                dmin = min(x["d"] for x in sel)
                if sum(1 for x in sel if x["d"] == dmin) != 1:
                    continue
            g.writelines(rec)
            n += 1
    assert n > 100
    return sub


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/krepp not built")
def test_place_on_a_query_tree_equals_the_reference(tmp_path):
    """`place -t NWK` (ref src/krepp.cpp:48-64, src/phytree.cpp:421-473): a tree with another topology, a multifurcating root, a
    leaf the index does not have (its parent is then no placement candidate and weighs its one covered child by 1) and without
    one indexed reference (dropped at colour expansion).  Against the reference CLI, on the reads whose closest reference -- among
    the references the tree keeps -- is unique; with the index's own tree as -t the output must not change at all."""
    idx, q = os.path.join(SMALL, "index"), os.path.join(SMALL, "reads.fq")
    own = run(EXE, "place", "--tabular", "-i", idx, "-q", q).splitlines()[3:]
    same = run(EXE, "place", "--tabular", "-i", idx, "-q", q, "-t", os.path.join(SMALL, "tree.nwk")).splitlines()[3:]
    assert own == same and len(own) == 408
    alt = tmp_path / "alt.nwk"
    alt.write_text(ALT_TREE + "\n")
    kept = {"G000000", "G000001", "G000002", "G000003", "G000004", "G000005", "G000006"}
    sub = untied_subset(tmp_path, kept)
    for extra in ([], ["--no-filter"], ["--no-multi"]):
        a = run(EXE, "place", "--tabular", "-i", idx, "-q", str(sub), "-t", str(alt), *extra).splitlines()
        b = run(REF, "place", "--tabular", "-i", idx, "-q", str(sub), "-t", str(alt), *extra).splitlines()
        assert a[1] == b[1] and "GXXXXXX" in a[1]                      # the edge-numbered query tree
        assert sorted(a[3:]) == sorted(b[3:]) and len(a) > 100, (extra, [x for x in a[3:] if x not in set(b)][:4], [x for x in b[3:] if x not in set(a)][:4])
    ja = rows(run(EXE, "place", "-i", idx, "-q", str(sub), "-t", str(alt)))
    jb = rows(run(REF, "place", "-i", idx, "-q", str(sub), "-t", str(alt)))
    assert ja == jb and len(ja) > 100
    d = run(EXE, "dist", "-i", idx, "-q", q).splitlines()[2:]          # dist takes no tree
    assert len(d) > 500


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/krepp not built")
def test_place_on_lineages_equals_the_reference(tmp_path):
    """`place -l FILE` (ref src/krepp.cpp:37-46,742-744, src/phytree.cpp:320-370): the taxonomy of a lineage file as the placement
    tree -- chains of one-child nodes (never candidates), no branch lengths (pendant and distal print as 0), a reference the
    index lacks, an indexed reference left out.  Against the reference CLI in every output form."""
    from lineages import KEPT, LINEAGES
    idx = os.path.join(SMALL, "index")
    lin = tmp_path / "lin.tsv"
    lin.write_text(LINEAGES)
    sub = untied_subset(tmp_path, KEPT)
    for extra in ([], ["--no-filter"], ["--no-multi"]):
        a = run(EXE, "place", "--tabular", "-i", idx, "-q", str(sub), "-l", str(lin), *extra).splitlines()
        b = run(REF, "place", "--tabular", "-i", idx, "-q", str(sub), "-l", str(lin), *extra).splitlines()
        assert a[1] == b[1] and "Bacteria{30})root{31};" in a[1]
        assert sorted(a[3:]) == sorted(b[3:]) and len(a) > 100, (extra, [x for x in a[3:] if x not in set(b)][:4], [x for x in b[3:] if x not in set(a)][:4])
        assert extra == ["--no-multi"] or any("\tGd\t19\t" in x or "\tPb\t29\t" in x for x in a[3:])    # placements on taxa
    ja = rows(run(EXE, "place", "-i", idx, "-q", str(sub), "-l", str(lin)))
    jb = rows(run(REF, "place", "-i", idx, "-q", str(sub), "-l", str(lin)))
    assert ja == jb and len(ja) > 100
    sa = run(EXE, "place", "--summarize", "-i", idx, "-q", str(sub), "-l", str(lin)).splitlines()
    sb = run(REF, "place", "--summarize", "-i", idx, "-q", str(sub), "-l", str(lin)).splitlines()
    assert sa[1] == sb[1] and len(sa) == len(sb)


@needs_shim
@pytest.mark.parametrize("with_tree,pre", [(True, ()), (False, ("--seed", "5"))], ids=["guide_tree", "no_tree_seeded"])
def test_shim_index_equals_the_reference_library(with_tree, pre, tmp_path):
    """`krepp_gpu index`: the reference's own driver (CLI11, LSH position draw, input map, guide tree, RSeq / kseq reader) with
    build_index / save_index redirected to the library builder (oracle/shim/gpubuilder.hpp).  The library must be the stock
    binary's up to the numbering of colours above the tree nodes, and the stock `krepp dist` must answer alike from either."""
    import shutil
    from libraries import assert_same_library, read_library
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/krepp not built")
    w = str(tmp_path)
    shutil.copytree(os.path.join(SMALL, "genomes"), os.path.join(w, "genomes"))
    for f in ("input_map.tsv", "tree.nwk"):
        shutil.copy(os.path.join(SMALL, f), w)
    args = ["-k", "23", "-w", "29", "-h", "8", "-m", "3", "-r", "1", "-i", "input_map.tsv", *(["-t", "tree.nwk"] if with_tree else [])]
    for exe, out in ((SHIM, "gpu_index"), (REF, "ref_index")):
        r = subprocess.run([exe, *pre, "index", *args, "-o", out], cwd=w, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    assert_same_library(read_library(os.path.join(w, "gpu_index")), read_library(os.path.join(w, "ref_index")))
    q = os.path.join(SMALL, "reads.fq")
    a = sorted(run(REF, "dist", "-i", os.path.join(w, "gpu_index"), "-q", q).splitlines()[2:])
    b = sorted(run(REF, "dist", "-i", os.path.join(w, "ref_index"), "-q", q).splitlines()[2:])
    assert a == b and len(a) > 100
