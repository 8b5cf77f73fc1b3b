#!/usr/bin/env python3
"""oracle/shim/patch_krepp.py REF_KREPP_CPP OUT_CPP -- TEST INFRASTRUCTURE.
Writes a copy of the reference's src/krepp.cpp (into oracle/_ref/, git-ignored) in which the two places that create an IBatch
(src/krepp.cpp:367 in estimate_distances, :460 in place_sequences) create the GpuBatch of oracle/shim/gpubatch.hpp instead.
`krepp index` is redirected too: the calls of IndexMultiple::build_index and save_index in main (src/krepp.cpp:729,732) become
gpu_build_index / gpu_save_index of oracle/shim/gpubuilder.hpp.  Nothing else changes; the script fails if the reference no
longer has exactly those calls."""
import sys

src = open(sys.argv[1]).read()
call = "std::make_shared<IBatch>(index, qs, hdist_th, chisq_value, dist_max, tau, no_filter, multi, summarize)"
assert src.count(call) == 2, "the reference's IBatch construction sites changed"
first = src.index(call)
src = src[:first] + "std::make_shared<GpuBatch>(gpu_index(index_dir), index, qs, hdist_th, chisq_value, dist_max, tau, no_filter, multi, summarize, false)" + src[first + len(call):]
second = src.index(call)
src = src[:second] + "std::make_shared<GpuBatch>(gpu_index(index_dir), index, qs, hdist_th, chisq_value, dist_max, tau, no_filter, multi, summarize, true)" + src[second + len(call):]
for call, repl in (("krepp_index.build_index();", "gpu_build_index(krepp_index);"), ("krepp_index.save_index();", "gpu_save_index(krepp_index);")):
    assert src.count(call) == 1, "the reference's index driver changed"
    src = src.replace(call, repl)
inc = '#include "krepp.hpp"'
assert src.count(inc) == 1
# GpuBatch takes the batch out of QSeq the way IBatch does (src/query.cpp:32-33), which QSeq allows to its friend IBatch only
# (src/rqseq.hpp:132-137).  A maintainer would add `friend class GpuBatch;` there; this build leaves the reference's headers
# alone and opens the access specifiers for this one translation unit instead (as oracle/ref_dump.cpp does).
src = src.replace(inc, '#include <bits/stdc++.h>\n#include <omp.h>\n#include <zlib.h>\n#define private public\n#define protected public\n' + inc +
                  '\n#undef private\n#undef protected\n#include "gpubatch.hpp"\n#include "gpubuilder.hpp"')
open(sys.argv[2], "w").write(src)
