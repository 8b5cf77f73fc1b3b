"""CPU: the tree `place -l` builds from a lineage file (HostTree::parse_lineages) against the one the UNMODIFIED reference
prints for the same file (ref src/phytree.cpp:320-370, src/krepp.cpp:37-46), and the reference's two error exits."""
import os
import subprocess

import pytest

from conftest import GOLDEN_DIR, REF_DIR, needs_ref
from lineages import LINEAGES

S = os.path.join(GOLDEN_DIR, "small")


def lineage_tree(path):
    import krepp_b200
    from krepp_b200 import capi
    ix = krepp_b200.Index(os.path.join(S, "index"), capi.DEVICE_NONE, lineages=str(path))
    try:
        return ix.jplace_tree()
    finally:
        ix.close()


@needs_ref
def test_lineage_tree_equals_the_reference(tmp_path):
    f = tmp_path / "lin.tsv"
    f.write_text(LINEAGES)
    one = tmp_path / "one.fq"
    with open(os.path.join(S, "reads.fq")) as g:
        one.write_text("".join(g.readline() for _ in range(4)))
    ref = subprocess.run([os.path.join(REF_DIR, "krepp"), "place", "--tabular", "-i", os.path.join(S, "index"), "-q", str(one), "-l", str(f)], capture_output=True,
                         text=True, check=True).stdout.splitlines()
    mine = lineage_tree(f)
    assert ref[1] == "# " + mine
    assert "GXXXXXX{17}" in mine and "(G000003{11})Gc{12})Fb{13}" in mine and mine.endswith(")Bacteria{30})root{31};")


@pytest.mark.parametrize("text,msg", [
    ("G000000\td__A;p__B\nG000000\td__A;p__C\n", "The same reference appears more than once in the lineage file."),
    ("G000000\td__A;p__B\nG000001\n", "Failed to reference to lineage mapping!"),
    ("G000000\td__A;p__B\n\nG000001\td__A\n", "Failed to reference to lineage mapping!"),
    ("G000000\td__A;p__B\nG000001\t\n", "Failed to reference to lineage mapping!"),
])
def test_lineage_errors_are_the_reference_s(tmp_path, text, msg):
    from krepp_b200.capi import KreppError
    f = tmp_path / "bad.tsv"
    f.write_text(text)
    with pytest.raises(KreppError, match=msg.replace(".", r"\.")):
        lineage_tree(f)
    exe = os.path.join(REF_DIR, "krepp")
    if os.path.exists(exe):
        r = subprocess.run([exe, "place", "-i", os.path.join(S, "index"), "-q", os.path.join(S, "reads.fq"), "-l", str(f)], capture_output=True, text=True)
        assert r.returncode != 0 and msg in r.stderr + r.stdout


def test_lineage_tree_shapes(tmp_path):
    """No taxa at all (the reference hangs below the root), the same word at two ranks (one node, its first parent), the rank
    prefix removed wherever "<character>__" occurs, a last line without a newline."""
    f = tmp_path / "l.tsv"
    f.write_text("A\t\tx\nB\td__X; p__X; c__Y\nC\td__X;c__q__Z")
    assert lineage_tree(f) == "(A{0},((B{1})Y{2},(C{3})Z{4})X{5})root{6};"
