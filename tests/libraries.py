"""Readers and checkers for the files `krepp index` writes, shared by the CPU test of the library
writer and the GPU test of the whole builder.  A library built here must be the reference's up to the numbering of the colours
above the tree nodes: same metadata, offsets, encodings, reference list, tree and rho, and every k-mer's colour expanding to the
same references."""
import os
import struct

import numpy as np


def suffix_of(index_dir: str) -> str:
    sfx = sorted(f[len("metadata"):] for f in os.listdir(index_dir) if f.startswith("metadata-") and not f.endswith(".txt"))
    assert len(sfx) == 1, sfx
    return sfx[0]


def read_library(index_dir: str) -> dict:
    sfx = suffix_of(index_dir)
    rd = lambda name: open(os.path.join(index_dir, name + sfx), "rb").read()
    md = rd("metadata")
    k, w, h = md[0], md[1], md[2]
    m, r = struct.unpack_from("<II", md, 3)
    out = dict(sfx=sfx, metadata=md, k=k, w=w, h=h, m=m, r=r, frac=md[11], nrows=struct.unpack_from("<I", md, 12)[0], ppos=bytes(md[16:16 + h]), npos=bytes(md[16 + h:16 + k]))
    cm = rd("cmer")
    nk = struct.unpack_from("<Q", cm, 0)[0]
    pairs = np.frombuffer(cm, "<u4", 2 * nk, 8).reshape(nk, 2)
    assert len(cm) == 8 + 8 * nk
    out.update(nkmers=nk, enc=pairs[:, 0].copy(), se=pairs[:, 1].copy(), inc_bytes=rd("inc"))
    cr = rd("crecord")
    nnodes, nsub = struct.unpack_from("<II", cr, 0)
    out.update(nnodes=nnodes, nsubsets=nsub, pse=np.frombuffer(cr, "<u4", 2 * nsub, 8).reshape(nsub, 2).copy(), rho=np.frombuffer(cr, "<f8", nnodes, 8 + 8 * nsub).copy())
    assert len(cr) == 8 + 8 * nsub + 8 * nnodes
    out["reflist"] = rd("reflist")
    tp = os.path.join(index_dir, "tree" + sfx)
    out["tree"] = open(tp, "rb").read() if os.path.exists(tp) else None
    return out


def colour_leaves(lib: dict) -> list:
    """colour id -> frozenset of the leaf node numbers it expands to (the walk of ref src/query.cpp:369-387: a tree node that is a
    leaf is itself, anything else splits through its pair, 0 is nothing)."""
    pse, nnodes = lib["pse"], lib["nnodes"]
    memo = [None] * len(pse)
    memo[0] = frozenset()
    for start in range(1, len(pse)):
        stack = [start]
        while stack:
            c = stack[-1]
            if memo[c] is not None:
                stack.pop()
                continue
            a, b = int(pse[c][0]), int(pse[c][1])
            if c < nnodes and a == 0 and b == c:
                memo[c] = frozenset([c])
                stack.pop()
                continue
            assert a != c and b != c and a < len(pse) and b < len(pse), (c, a, b)
            todo = [x for x in (a, b) if memo[x] is None]
            if todo:
                stack.extend(todo)
                continue
            memo[c] = memo[a] | memo[b]
            stack.pop()
    return memo


def assert_same_library(mine: dict, ref: dict):
    for f in ("metadata", "inc_bytes", "reflist", "tree"):
        assert mine[f] == ref[f], f
    assert mine["nkmers"] == ref["nkmers"] and (mine["enc"] == ref["enc"]).all()
    assert mine["nnodes"] == ref["nnodes"]
    assert mine["rho"].tobytes() == ref["rho"].tobytes()
    # the tree's own colours are numbered alike
    n = mine["nnodes"]
    a, b = colour_leaves(mine), colour_leaves(ref)
    assert a[:n] == b[:n]
    sa, sb = mine["se"], ref["se"]
    assert all(a[int(x)] == b[int(y)] for x, y in zip(sa, sb))
    return len(mine["pse"]), len(ref["pse"])


def numpy_union(tables: dict) -> tuple:
    """What krepp_builder_union computes, restated with numpy: tables = {leaf rank: sorted unique keys}.  Returns
    (keys, set_of, set_begin, set_leaves) with the sets numbered by first appearance."""
    ranks = sorted(tables)
    keys = np.concatenate([tables[r] for r in ranks]) if ranks else np.zeros(0, np.uint64)
    leaf = np.concatenate([np.full(len(tables[r]), r, np.uint32) for r in ranks]) if ranks else np.zeros(0, np.uint32)
    order = np.lexsort((leaf, keys))
    keys, leaf = keys[order], leaf[order]
    head = np.ones(len(keys), bool)
    head[1:] = keys[1:] != keys[:-1]
    starts = np.flatnonzero(head)
    ends = np.append(starts[1:], len(keys))
    ids, set_of, sets = {}, np.zeros(len(starts), np.uint32), []
    for i, (s, e) in enumerate(zip(starts, ends)):
        t = tuple(int(x) for x in leaf[s:e])
        if t not in ids:
            ids[t] = len(sets)
            sets.append(t)
        set_of[i] = ids[t]
    set_begin = np.zeros(len(sets) + 1, np.uint64)
    set_begin[1:] = np.cumsum([len(t) for t in sets])
    set_leaves = np.array([x for t in sets for x in t], np.uint32)
    return keys[starts].copy(), set_of, set_begin, set_leaves


def same_colours_vectorised(mine: dict, ref: dict) -> bool:
    """assert_same_library's per-k-mer colour check for libraries of millions of k-mers: colour ids of both libraries are mapped
    to one shared numbering of their expansions, then the two columns are compared as arrays."""
    shared = {}
    canon = []
    for lib in (mine, ref):
        exp = colour_leaves(lib)
        canon.append(np.array([shared.setdefault(s, len(shared)) for s in exp], np.int64))
    return bool((canon[0][mine["se"]] == canon[1][ref["se"]]).all())


def sampled_colour_check(mine: dict, ref: dict, n: int = 200_000, seed: int = 1) -> tuple:
    """For libraries too large to expand every colour in Python: n random k-mers, their colours expanded on demand in both
    libraries.  Returns (k-mers checked, k-mers whose two colours expand to different references)."""
    rng = np.random.default_rng(seed)
    at = rng.integers(0, len(mine["se"]), min(n, len(mine["se"])))

    def expander(lib):
        pse, nnodes, memo = lib["pse"], lib["nnodes"], {0: frozenset()}

        def expand(c):
            stack = [c]
            while stack:
                x = stack[-1]
                if x in memo:
                    stack.pop()
                    continue
                a, b = int(pse[x][0]), int(pse[x][1])
                if x < nnodes and a == 0 and b == x:
                    memo[x] = frozenset([x])
                    stack.pop()
                    continue
                todo = [y for y in (a, b) if y not in memo]
                if todo:
                    stack.extend(todo)
                    continue
                memo[x] = memo[a] | memo[b]
                stack.pop()
            return memo[c]
        return expand

    ea, eb = expander(mine), expander(ref)
    bad = sum(1 for i in at if ea(int(mine["se"][i])) != eb(int(ref["se"][i])))
    return len(at), bad
