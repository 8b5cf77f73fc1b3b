#!/bin/bash
# compute-sanitizer over this round's new kernels on the small fixtures: memcheck on the cut-read path (segment_combine_kernel),
# seek (sketch image + seek_kernel), the sketch builder (minimizer kernel with its HyperLogLog registers) and the lineage / query
# tree placements; racecheck on the kernels that share memory inside a warp or CTA (segment_combine, place_collect).
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_sanitize2.sh <tag>'
TAG=${1:-san2}; O=gpurun_out/$TAG; mkdir -p $O
SAN=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
T1="tests/test_gpu_segments.py::test_cut_reads_equal_oracle tests/test_gpu_seek.py::test_seek_equals_oracle_and_reference tests/test_gpu_sketch.py::test_sketch_of_contigs_seeded_and_gzip tests/test_gpu_minimizer.py"
timeout 700 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -x -q -m gpu $T1 -k "not k28 and not k29" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"
grep -n "ERROR SUMMARY\|passed\|failed" $O/memcheck.log | head
timeout 500 $SAN --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest -x -q -m gpu "tests/test_gpu_segments.py::test_cut_reads_equal_uncut_reads_bit_for_bit" > $O/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -n "RACECHECK SUMMARY\|passed\|failed\|hazard" $O/racecheck.log | head
