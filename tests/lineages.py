"""A Greengenes/GTDB style lineage file over the golden genomes for `place -l` (ref src/phytree.cpp:320-370): one top-level
taxon (with several the reference hangs them below the root in the order of a hash map), an empty rank ("s__"), a rank shared
by two references, a reference the index does not have, one indexed reference left out, a trailing ignored column."""
LINEAGES = """\
G000000\td__Bacteria; p__Pa; c__Ca; o__Oa; f__Fa; g__Ga; s__Ga one
G000001\td__Bacteria; p__Pa; c__Ca; o__Oa; f__Fa; g__Ga; s__Ga one
G000002\td__Bacteria; p__Pa; c__Ca; o__Oa; f__Fa; g__Gb; s__Gb two
G000003\td__Bacteria; p__Pb; c__Cb; o__Ob; f__Fb; g__Gc; s__
G000004\td__Bacteria;p__Pb;c__Cb;o__Ob;f__Fc;g__Gd;s__Gd three
G000005\td__Bacteria; p__Pb; c__Cb; o__Ob; f__Fc; g__Gd; s__Gd four
GXXXXXX\td__Bacteria; p__Pb; c__Cb; o__Ob; f__Fc; g__Gd; s__Gd four
G000006\td__Bacteria; p__Pb; c__Cc; o__Oc; f__Fd; g__Ge; s__Ge five\textra column
"""
KEPT = {"G000000", "G000001", "G000002", "G000003", "G000004", "G000005", "G000006"}
