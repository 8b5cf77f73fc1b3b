#!/bin/bash
# compute-sanitizer over the small-fixture GPU tests: memcheck (golden index: dist, place, sorted chain, shards, brief rows,
# index-geometry variants) and racecheck (shared-memory hazards of the fused kernel, the join, the resolve sort, placement).
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_sanitize.sh <tag>'
TAG=${1:-san}; O=gpurun_out/$TAG; mkdir -p $O
SAN=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
SMALL="tests/test_gpu_parity.py::test_golden_small_index_k21_h7 tests/test_gpu_sorted.py::test_sorted_golden_small_index tests/test_gpu_parity.py::test_brief_rows_equal_full_rows tests/test_gpu_shard.py::test_small_index_place_through_shards tests/test_gpu_shard.py::test_shard_edge_cases"
timeout 1000 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -x -q $SMALL tests/test_gpu_variants.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?"
grep -n "ERROR SUMMARY\|passed\|failed" $O/memcheck.log | head
timeout 1200 $SAN --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest -x -q $SMALL > $O/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -n "RACECHECK SUMMARY\|passed\|failed\|hazard" $O/racecheck.log | head
