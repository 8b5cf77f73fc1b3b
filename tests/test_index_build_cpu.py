"""CPU: the host stage of `krepp index` -- krepp_builder_set_union / krepp_builder_write: reference sets ->
colours along the guide tree -> the library files -- against libraries built by the UNMODIFIED reference (`oracle/_ref/krepp
index`) from the same genomes.  The union handed in is computed here from the oracle's leaf tables (test plumbing; on the GPU box
krepp_builder_add_genome / krepp_builder_union compute it, tests/test_gpu_index_build.py).  The library written must be the
reference's up to the numbering of the colours above the tree nodes, and the reference's own `krepp dist` must print the same
lines from either."""
import os
import subprocess

import numpy as np
import pytest

import conftest
import krepp_b200
import oracle_lib as O
from conftest import needs_ref
from libraries import assert_same_library, colour_leaves, numpy_union, read_library
from test_seek_cpu import fasta_seqs
from variants import SMALL, TREELESS, VARIANTS, build

REF = os.path.join(conftest.REF_DIR, "krepp")
NONE = -1  # KREPP_DEVICE_NONE


def names_and_paths():
    rows = [l.rstrip("\n").split("\t") for l in open(os.path.join(SMALL, "input_map.tsv")) if l.strip()]
    return [r[0] for r in rows], {r[0]: os.path.join(SMALL, r[1]) for r in rows}


def host_build(ref_lib: dict, out_dir: str, nwk: str | None, names, paths, seed=None):
    """The library of the golden genomes under the geometry of `ref_lib`, written by krepp_builder_write from a numpy union of
    the oracle's leaf tables."""
    g = krepp_b200.Index.geometry(ref_lib["k"], ref_lib["w"], ref_lib["h"], ref_lib["m"], ref_lib["r"], bool(ref_lib["frac"]), seed=seed, device=NONE)
    b = krepp_b200.LibraryBuilder(g, nwk, names)
    tables, rho = {}, np.zeros(b.nleaves)
    for nm in names:
        rank = b.leaf_rank(nm)
        if rank is None:
            continue
        keys, rh = O.oracle_sketch_table(ref_lib, fasta_seqs(paths[nm]))
        tables[rank], rho[rank] = keys, rh
    keys, set_of, set_begin, set_leaves = numpy_union(tables)
    b.set_union(keys, set_of, set_begin, set_leaves, rho)
    nk, nsub = b.write(out_dir)
    b.close()
    g.close()
    return nk, nsub


def ref_dist(index_dir, extra=()):
    out = subprocess.run([REF, "dist", "-i", index_dir, "-q", os.path.join(SMALL, "reads.fq"), *extra], capture_output=True, text=True, check=True).stdout
    return sorted(l for l in out.splitlines() if not l.startswith("#"))


@needs_ref
@pytest.mark.parametrize("label,args", [("default_k21", ["-k", "21", "-w", "25", "-h", "7"])] + VARIANTS[:3], ids=lambda v: v if isinstance(v, str) else "")
def test_written_library_equals_the_reference(label, args, tmp_path_factory, tmp_path):
    ref_dir = build("lib_" + label, args, tmp_path_factory.getbasetemp())
    ref = read_library(ref_dir)
    names, paths = names_and_paths()
    nwk = open(os.path.join(SMALL, "tree.nwk")).read()
    mine_dir = str(tmp_path / "index")
    nk, nsub = host_build(ref, mine_dir, nwk, names, paths)
    mine = read_library(mine_dir)
    assert nk == ref["nkmers"] and nsub == mine["nsubsets"]
    n_mine, n_ref = assert_same_library(mine, ref)
    assert n_mine <= n_ref  # pairs are shared between sets; the reference's record may hold sets only met on the way up
    txt = open(os.path.join(mine_dir, "metadata" + mine["sfx"] + ".txt")).read()
    assert f"total_num_kmers: {nk}\n" in txt and f"nrows: {ref['nrows']}\n" in txt
    # the reference reads it and answers as it does from its own
    assert ref_dist(mine_dir) == ref_dist(ref_dir)
    assert ref_dist(mine_dir, ["--hdist-th", "3"]) == ref_dist(ref_dir, ["--hdist-th", "3"])  # (no --filter: its test is against a tie-dependent closest)
    # and so does this implementation's loader (host image only)
    a, b = krepp_b200.Index(mine_dir, device=NONE), krepp_b200.Index(ref_dir, device=NONE)
    assert a.info.nkmers == b.info.nkmers and a.info.nnodes == b.info.nnodes
    ca, cb = a.host_checksums(), b.host_checksums()
    assert ca[1] == cb[1]  # the offsets; the other sums weigh colour ids, which are numbered differently above the tree nodes
    a.close(); b.close()


@needs_ref
def test_written_library_without_a_guide_tree(tmp_path_factory, tmp_path):
    label, args = TREELESS
    ref_dir = build("lib_" + label, args, tmp_path_factory.getbasetemp(), with_tree=False)
    ref = read_library(ref_dir)
    assert ref["tree"] is None
    names, paths = names_and_paths()
    mine_dir = str(tmp_path / "index")
    host_build(ref, mine_dir, None, names, paths)
    mine = read_library(mine_dir)
    assert_same_library(mine, ref)
    assert ref_dist(mine_dir) == ref_dist(ref_dir)


@needs_ref
def test_written_library_of_draft_assemblies(tmp_path):
    """Contigs with runs of N, short contigs and the end-of-sequence emit; a leaf of the tree without a genome (an empty table, rho
    0) and a reference id the tree does not have (listed in reflist-*, never indexed) -- as the reference treats them."""
    import shutil
    from test_gpu_sketch import contigs_fasta
    w = str(tmp_path / "w")
    os.makedirs(w)
    shutil.copytree(os.path.join(SMALL, "genomes"), os.path.join(w, "genomes"))
    shutil.copy(os.path.join(SMALL, "tree.nwk"), w)
    contigs_fasta(os.path.join(w, "genomes", "G000001.fna"))
    names, paths = names_and_paths()
    names = ["GXXXXXX" if n == "G000005" else n for n in names]
    with open(os.path.join(w, "input_map.tsv"), "w") as f:
        for n in names:
            f.write(n + "\t./genomes/" + ("G000005" if n == "GXXXXXX" else n) + ".fna\n")
    subprocess.run([REF, "index", "-k", "21", "-w", "25", "-h", "7", "-o", "ref_index", "-i", "input_map.tsv", "-t", "tree.nwk"], cwd=w, check=True, capture_output=True)
    ref = read_library(os.path.join(w, "ref_index"))
    paths = {n: os.path.join(w, "genomes", ("G000005" if n == "GXXXXXX" else n) + ".fna") for n in names}
    mine_dir = os.path.join(w, "my_index")
    host_build(ref, mine_dir, open(os.path.join(w, "tree.nwk")).read(), names, paths)
    mine = read_library(mine_dir)
    assert_same_library(mine, ref)
    assert b"GXXXXXX\n" in mine["reflist"] and (mine["rho"] == 0).sum() == (ref["rho"] == 0).sum()
    assert ref_dist(mine_dir) == ref_dist(os.path.join(w, "ref_index"))


def random_sets(rng, nleaves, n_sets):
    sets = {tuple(range(nleaves))}
    while len(sets) < n_sets:
        k = int(rng.integers(1, nleaves + 1))
        sets.add(tuple(sorted(int(x) for x in rng.choice(nleaves, k, replace=False))))
    return sorted(sets)


@pytest.mark.parametrize("nwk", ["((A:1,B:1,C:1,D:1)x:1,(E:1,(F:1,G:1,H:1):1):1,I:1);", "(A,(B,(C,(D,(E,(F,(G,(H,I))))))));",
                                 "(((A,B),(C,D)),((E,F),(G,(H,I))));"], ids=["multifurcating", "caterpillar", "balanced"])
def test_colours_expand_to_their_sets_on_any_tree(nwk, tmp_path):
    """Every distinct reference set gets a colour that expands to exactly that set; a whole subtree is the tree node itself; pairs
    are shared.  (Synthetic sets; the reference's own multifurcating fixture is test_toy_library_equals_the_reference.)"""
    rng = np.random.default_rng(5)
    names = list("ABCDEFGHI")
    g = krepp_b200.Index.geometry(21, 25, 7, 4, 1, True, device=NONE)
    b = krepp_b200.LibraryBuilder(g, nwk, names)
    assert b.nleaves == 9 and b.leaf_rank("nope") is None
    sets = random_sets(rng, 9, 150)
    set_begin = np.zeros(len(sets) + 1, np.uint64)
    set_begin[1:] = np.cumsum([len(s) for s in sets])
    set_leaves = np.array([x for s in sets for x in s], np.uint32)
    keys = (np.arange(len(sets), dtype=np.uint64) * 7 + 3) << np.uint64(20)  # ascending, spread over rows
    b.set_union(keys, np.arange(len(sets), dtype=np.uint32), set_begin, set_leaves, np.linspace(0.1, 0.2, 9))
    out = str(tmp_path / "index")
    nk, nsub = b.write(out)
    lib = read_library(out)
    assert nk == len(sets) and nsub == lib["nsubsets"]
    exp = colour_leaves(lib)
    ix = krepp_b200.Index(out, device=NONE)  # leaf rank -> node number through this implementation's own tree
    leaf_se = [se for se in range(1, ix.info.nnodes + 1) if ix.node_name(se) in names]
    rank_of = {nm: b.leaf_rank(nm) for nm in names}
    se_of_rank = {rank_of[ix.node_name(se)]: se for se in leaf_se}
    for s, colour in zip(sets, lib["se"]):
        assert exp[int(colour)] == frozenset(se_of_rank[r] for r in s), s
    assert exp[ix.info.root_se] == frozenset(leaf_se)                    # the set of all references is the root
    assert int(lib["se"][sets.index(tuple(range(9)))]) == ix.info.root_se
    assert lib["rho"][se_of_rank[0]] == 0.1 and lib["rho"][0] == 0.0
    # shared pairs: far fewer colours than the sum of the set sizes
    assert lib["nsubsets"] < ix.info.nnodes + 1 + sum(len(s) - 1 for s in sets)
    ix.close(); b.close(); g.close()


def test_builder_argument_errors(tmp_path):
    g = krepp_b200.Index.geometry(21, 25, 7, 4, 1, True, device=NONE)
    with pytest.raises(krepp_b200.KreppError, match="two leaves named"):
        krepp_b200.LibraryBuilder(g, "((A,B),A);", ["A", "B"])
    with pytest.raises(krepp_b200.KreppError):
        krepp_b200.LibraryBuilder(g, "((A,B);", ["A", "B"])
    b = krepp_b200.LibraryBuilder(g, "((A,B),C);", ["A", "B", "C"])
    with pytest.raises(krepp_b200.KreppError, match="no union yet"):
        b.write(str(tmp_path / "x"))
    with pytest.raises(krepp_b200.KreppError, match="GPU"):  # no CPU fallback for the device stage
        b.add_genome("A", [b"ACGT" * 50])
    with pytest.raises(krepp_b200.KreppError, match="GPU"):
        b.union()
    b.set_union(np.array([5, 4], np.uint64), np.zeros(2, np.uint32), np.array([0, 1], np.uint64), np.array([0], np.uint32))
    with pytest.raises(krepp_b200.KreppError, match="ascending"):
        b.write(str(tmp_path / "x"))
    b.set_union(np.array([4, 5], np.uint64), np.zeros(2, np.uint32), np.array([0, 2], np.uint64), np.array([1, 0], np.uint32))
    with pytest.raises(krepp_b200.KreppError, match="ascending list of leaf ranks"):
        b.write(str(tmp_path / "x"))
    b.close(); g.close()


def test_geometry_with_given_positions():
    """krepp_geometry_open_positions: the masks of a drawn geometry come back when its positions are handed in (any order), and
    bad positions are refused."""
    drawn = krepp_b200.Index.geometry(27, 35, 11, 4, 1, True, device=NONE)
    ppos = [p for p in range(31, -1, -1) if (drawn.info.mask_hash_bp >> (2 * p)) & 3]
    assert ppos == [26, 24, 22, 21, 17, 14, 8, 7, 5, 3, 2]  # the reference's default draw
    given = krepp_b200.Index.geometry(27, 35, 11, 4, 1, True, device=NONE, ppos=bytes(reversed(ppos)))
    assert (given.info.mask_hash_bp, given.info.mask_drop_lr, given.info.nrows) == (drawn.info.mask_hash_bp, drawn.info.mask_drop_lr, drawn.info.nrows)
    other = krepp_b200.Index.geometry(27, 35, 11, 4, 1, True, device=NONE, ppos=bytes(range(11)))
    assert other.info.mask_hash_bp == (1 << 22) - 1
    for bad in (bytes([3] * 11), bytes(list(range(10)) + [27])):
        with pytest.raises(krepp_b200.KreppError, match="distinct and below k"):
            krepp_b200.Index.geometry(27, 35, 11, 4, 1, True, device=NONE, ppos=bad)
    for g in (drawn, given, other):
        g.close()


def random_newick(rng, names, max_children=4):
    """A random rooted tree over the names: leaves are joined in random groups of 2..max_children until one node is left."""
    nodes = list(names)
    rng.shuffle(nodes)
    while len(nodes) > 1:
        k = int(min(len(nodes), rng.integers(2, max_children + 1)))
        at = int(rng.integers(0, len(nodes) - k + 1))
        nodes[at:at + k] = ["(" + ",".join(nodes[at:at + k]) + ")"]
    return nodes[0] + ";"


@pytest.mark.parametrize("seed", range(6))
def test_colours_on_random_trees(seed, tmp_path):
    """Random trees with up to four children per node and random reference sets: every colour expands to exactly its set, the
    colour of a whole subtree is the subtree's node, equal sets share one id, and the record holds no colour that nothing reaches."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(5, 60))
    names = [f"L{i:02d}" for i in range(n)]
    nwk = random_newick(rng, names, max_children=2 if seed % 2 else 4)
    g = krepp_b200.Index.geometry(21, 25, 7, 4, 1, True, device=NONE)
    b = krepp_b200.LibraryBuilder(g, nwk, names)
    sets = random_sets(rng, n, 400)
    # clustered sets too (runs of neighbouring leaves: whole subtrees and near-subtrees), as real libraries have
    for _ in range(200):
        a = int(rng.integers(0, n)); z = int(rng.integers(a + 1, n + 1))
        s = set(range(a, z))
        if rng.random() < 0.5 and len(s) > 1:
            s.discard(int(rng.integers(a, z)))
        sets.append(tuple(sorted(s)))
    sets = sorted(set(sets))
    set_begin = np.zeros(len(sets) + 1, np.uint64)
    set_begin[1:] = np.cumsum([len(s) for s in sets])
    keys = np.arange(1, 2 * len(sets) + 1, dtype=np.uint64) << np.uint64(18)
    set_of = np.concatenate([np.arange(len(sets)), np.arange(len(sets))[::-1]]).astype(np.uint32)  # every set used by two k-mers
    b.set_union(keys, set_of, set_begin, np.array([x for s in sets for x in s], np.uint32), np.full(n, 0.2))
    out = str(tmp_path / "index")
    b.write(out)
    lib = read_library(out)
    exp = colour_leaves(lib)
    ix = krepp_b200.Index(out, device=NONE)
    se_of_rank = {b.leaf_rank(ix.node_name(se)): se for se in range(1, ix.info.nnodes + 1) if ix.node_name(se) in set(names)}
    assert len(se_of_rank) == n
    colour_of_set = {}
    for i, colour in zip(set_of, lib["se"]):
        want = frozenset(se_of_rank[r] for r in sets[int(i)])
        assert exp[int(colour)] == want
        assert colour_of_set.setdefault(int(i), int(colour)) == int(colour)
    # a set that is everything below a node is that node
    below = {exp[se]: se for se in range(1, ix.info.nnodes + 1)}
    for i, colour in colour_of_set.items():
        want = frozenset(se_of_rank[r] for r in sets[i])
        if want in below:
            assert colour == below[want]
    # every colour above the tree nodes is reachable from a k-mer's colour or from a node with more than two children
    reach, todo = set(), [int(c) for c in lib["se"]] + list(range(1, ix.info.nnodes + 1))
    while todo:
        c = todo.pop()
        if c in reach or c == 0:
            continue
        reach.add(c)
        a, z = int(lib["pse"][c][0]), int(lib["pse"][c][1])
        if not (a == 0 and z == c):
            todo += [a, z]
    assert reach == set(range(1, lib["nsubsets"]))
    ix.close(); b.close(); g.close()


TOY_TARBALL = "/root/reference/test/references_toy.tar.gz"


@needs_ref
@pytest.mark.skipif(not os.path.exists(TOY_TARBALL), reason="the reference's toy genomes are only in this container")
def test_toy_library_equals_the_reference(tmp_path):
    """Configuration 1, the reference's own fixture: 25 genomes (71 Mbp, draft assemblies of thousands of contigs) on a guide tree
    with multifurcations.  The library written from the oracle's leaf tables has the reference-built toy index's metadata, offsets,
    encodings and rho byte for byte, and each of its 6.9 M k-mers expands to the same references.  (The reference's colours of the
    multifurcating nodes themselves expand to their first child only -- SURVEY.md 7.7 -- but no k-mer of its index carries them: a
    k-mer held by every child gets a separate colour there, the node itself here.)"""
    from libraries import same_colours_vectorised
    toy = conftest.TOY_DIR
    subprocess.run(["tar", "-C", str(tmp_path), "-xzf", TOY_TARBALL], check=True)
    subprocess.run("xz -d -f " + str(tmp_path / "references_toy") + "/*.xz", shell=True, check=True)
    ref = read_library(os.path.join(toy, "index_toy"))
    rows = [l.rstrip("\n").split("\t") for l in open(os.path.join(toy, "input_map.tsv")) if l.strip()]
    names, paths = [r[0] for r in rows], {r[0]: str(tmp_path / r[1]) for r in rows}
    mine_dir = str(tmp_path / "index")
    nk, nsub = host_build(ref, mine_dir, open(os.path.join(toy, "tree_toy.nwk")).read(), names, paths)
    mine = read_library(mine_dir)
    for f in ("metadata", "inc_bytes", "reflist", "tree"):
        assert mine[f] == ref[f], f
    assert nk == ref["nkmers"] == 6934548 and (mine["enc"] == ref["enc"]).all() and mine["rho"].tobytes() == ref["rho"].tobytes()
    assert same_colours_vectorised(mine, ref)
    a, b = colour_leaves(mine), colour_leaves(ref)
    leaves = [se for se in range(1, ref["nnodes"]) if a[se] == frozenset([se])]
    assert len(leaves) == 25 and a[ref["nnodes"] - 1] == frozenset(leaves)     # here the root is every reference ...
    assert b[ref["nnodes"] - 1] < frozenset(leaves)                             # ... in the reference's record it is not


@needs_ref
def test_partial_libraries_written_into_one_directory(tmp_path_factory, tmp_path):
    """Three builds with one LSH residue each (m = 4, r = 0, 2, 3, --no-frac) into ONE directory, as `krepp index` is run for a
    library split by residue (ref src/krepp.cpp:66-108): every partial equals the reference's, the reference's `dist` reads the
    directory alike, and this implementation's loader merges it to the same image as the reference-built directory."""
    from variants import build_partials
    ref_dir = build_partials(tmp_path_factory.getbasetemp(), rs=("0", "2", "3"), frac=False)
    names, paths = names_and_paths()
    nwk = open(os.path.join(SMALL, "tree.nwk")).read()
    mine_dir = str(tmp_path / "index")
    for r in (0, 2, 3):
        sfx = f"-m4r{r}-no_frac"
        md = open(os.path.join(ref_dir, "metadata" + sfx), "rb").read()
        geom = dict(k=md[0], w=md[1], h=md[2], m=4, r=r, frac=0, ppos=bytes(md[16:16 + md[2]]))
        host_build(geom, mine_dir, nwk, names, paths)
        for f in ("metadata", "inc", "reflist", "tree"):
            assert open(os.path.join(mine_dir, f + sfx), "rb").read() == open(os.path.join(ref_dir, f + sfx), "rb").read(), (f, r)
    assert sorted(os.listdir(mine_dir)) == sorted(os.listdir(ref_dir))
    assert ref_dist(mine_dir) == ref_dist(ref_dir)
    a, b = krepp_b200.Index(mine_dir, device=NONE), krepp_b200.Index(ref_dir, device=NONE)
    assert a.info.nkmers == b.info.nkmers and a.info.nrows == b.info.nrows and a.host_checksums()[1] == b.host_checksums()[1]
    a.close(); b.close()
