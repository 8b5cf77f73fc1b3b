"""CPU tests of the C-ABI library: it loads, exports every symbol include/krepp_b200.h declares, parses the on-disk
index exactly like the oracle (host logic), and fails loudly -- not silently -- when asked to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT, TOY_DIR, needs_ref

SMALL = os.path.join(GOLDEN_DIR, "small", "index")


@pytest.fixture(scope="module")
def lib():
    import krepp_b200
    krepp_b200.build_library()
    return krepp_b200.load_library()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "krepp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(krepp_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 18
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} is declared in include/krepp_b200.h but not exported"
    assert lib.krepp_abi_version() == 2


def test_header_is_plain_c(tmp_path):
    """The boundary header compiles as C11 with no C++ or torch types in the signatures."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "krepp_b200.h"\nint main(void){ krepp_params_t p; (void)p; return sizeof(krepp_record_t) == 56 && sizeof(krepp_read_summary_t) == 44 && sizeof(krepp_placement_t) == 56 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_struct_layouts_match_numpy_views():
    from krepp_b200 import capi
    assert capi.RECORD_DTYPE.itemsize == 56 and capi.READ_DTYPE.itemsize == 44 and capi.PLACEMENT_DTYPE.itemsize == 56


def test_host_index_matches_oracle(lib):
    """Index parsing on the host (metadata, masks, tree numbering, names, edge-numbered Newick) against the oracle."""
    import krepp_b200
    import oracle_lib as O
    for d in [SMALL] + ([os.path.join(TOY_DIR, "index_toy")] if os.path.isdir(os.path.join(TOY_DIR, "index_toy")) else []):
        ix, o = krepp_b200.Index(d, device=-1), O.OracleIndex(d)
        i = ix.info
        assert (i.k, i.h, i.m, i.nnodes) == (o.k, o.hh, o.m, o.nnodes)
        assert i.mask_hash_bp == O.lib().ko_index_mask_hash_bp(o.h) and i.mask_drop_lr == O.lib().ko_index_mask_drop_lr(o.h)
        assert ix.jplace_tree() == o.jplace_tree()
        t = ix.tree()
        for se in range(1, o.nnodes + 1):
            assert ix.node_name(se) == o.name(se)
            assert t["parent"][se] == O.lib().ko_index_parent(o.h, se)
            assert bool(t["is_leaf"][se]) == bool(O.lib().ko_index_is_leaf(o.h, se))
            b = O.lib().ko_index_blen(o.h, se)
            assert (np.isnan(t["blen"][se]) and np.isnan(b)) or t["blen"][se] == b
        assert i.nleaves == int(t["is_leaf"].sum()) and i.root_se == o.nnodes
        ix.close()


def test_newick_conventions(tmp_path):
    """Reference Newick conventions (src/phytree.cpp:84-215): post-order se, unlabeled internal nodes named se-1 / NA,
    quotes, missing lengths -> distal 0, unifurcation and trailing-garbage errors."""
    import shutil
    import krepp_b200
    from krepp_b200.capi import KreppError

    def with_tree(nwk):
        d = tmp_path / f"ix{abs(hash(nwk))}"
        shutil.copytree(SMALL, d)
        (d / "tree-m4r1-frac").write_text(nwk)
        return str(d)
    names = ["G%06d" % i for i in range(8)]
    # same topology as the golden tree (the colour record is tied to it), different labels / lengths / quoting
    nwk = "(((%s:0.1,%s:0.2):0.05,%s)X:0.1,('%s':0.3,((%s:1,%s:2)Y,(%s:1e-3,%s:0.5):0.25):0.125):0.5)root:0;\n" % tuple(names)
    ix = krepp_b200.Index(with_tree(nwk), device=-1)
    t = ix.tree()
    assert ix.info.nnodes == 15 and ix.info.nleaves == 8 and ix.info.root_se == 15
    assert [ix.node_name(s) for s in (1, 2, 3, 4, 5, 6, 9)] == [names[0], names[1], "2", names[2], "X", names[3], "Y"]
    assert ix.node_name(3, True) == "NA" and t["nchildren"][15] == 2 and t["parent"][3] == 5 and t["parent"][5] == 15
    assert np.isnan(t["blen"][4]) and t["blen"][6] == 0.3 and np.isnan(t["blen"][9]) and t["blen"][12] == 0.25
    assert ix.jplace_tree().startswith("(((G000000:0.10000{0},G000001:0.20000{1}):0.05000{2},G000002{3})X:0.10000{4},(G000003:0.30000{5},")
    ix.close()
    for bad, msg in [("((A,B),(C));", "single child"), ("(A,B)", "other than ';'"), ("(A,B);(C,D);", "';'"), ("(A[x],B);", "'[' or ']'")]:
        with pytest.raises(KreppError) as e:
            krepp_b200.Index(with_tree(bad), device=-1)
        assert e.value.code == 2 and msg in str(e.value)


def test_errors_are_loud(lib, tmp_path):
    import krepp_b200
    from krepp_b200.capi import KreppError
    with pytest.raises(KreppError) as e:
        krepp_b200.Index(str(tmp_path), device=-1)
    assert e.value.code == 2  # KREPP_ERR_IO
    ix = krepp_b200.Index(SMALL, device=-1)
    with pytest.raises(KreppError) as e:
        krepp_b200.IBatch(ix, [b"ACGT" * 10])
    assert e.value.code == 3 and "GPU" in str(e.value)  # KREPP_ERR_CUDA: no CPU fallback
    with pytest.raises(KreppError) as e:
        krepp_b200.IBatch(ix, [b"ACGT" * 10], place=True, hdist_th=1, tau=2)
    assert e.value.code == 1 and "tau" in str(e.value)  # src/krepp.hpp:192-199
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(KreppError) as e:
            krepp_b200.Index(SMALL, device=0)
        assert e.value.code == 3
    # truncated index files are rejected with the reference's wording
    import shutil
    d = tmp_path / "trunc"
    shutil.copytree(SMALL, d)
    with open(d / "cmer-m4r1-frac", "r+b") as f:
        f.truncate(1000)
    with pytest.raises(KreppError) as e:
        krepp_b200.Index(str(d), device=-1)
    assert "k-mer vector" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """The product path must not import, link or execute anything under oracle/ (that would void every parity claim)."""
    import subprocess
    for root, _, files in os.walk(os.path.join(ROOT, "krepp_b200")):
        if "_build" in root:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(root, fn), errors="ignore").read()
                assert "oracle" not in txt.lower() or fn == "__init__.py", f"{fn} mentions the oracle"
    so = os.path.join(ROOT, "krepp_b200", "_build", "libkrepp_b200.so")
    out = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    assert " ko_" not in syms
