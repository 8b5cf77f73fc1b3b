#!/bin/bash
# First GPU contact of a sorted-pipeline change: one small parity test under compute-sanitizer, the sorted-pipeline parity
# tests, then config-3 timings of both pipelines (parity sample against the oracle included).
# usage: gpurun --timeout 1800 -- 'bash tools/gpu_sorted.sh <tag> [reads] [full]'
TAG=${1:-s}; READS=${2:-2000000}; O=gpurun_out/$TAG; mkdir -p $O
( KREPP_PIPELINE=sorted timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sorted.py -x -q -k golden ) > $O/sanitizer.log 2>&1; echo "sanitizer rc=$?"
grep -v "^=========     \|^$" $O/sanitizer.log | tail -25
( time timeout 1500 python -m pytest tests/test_gpu_sorted.py -x -q ) > $O/pytest_sorted.log 2>&1; echo "pytest sorted rc=$?" | tee -a $O/pytest_sorted.log
tail -25 $O/pytest_sorted.log
timeout 1200 python tools/perf_c3.py --reads $READS --batch 1000000 --check 300 --skip-cli --cpu-reads 0 > $O/perf_c3.log 2>&1; echo "perf_c3 rc=$?"
tail -12 $O/perf_c3.log
if [ "$3" == "full" ]; then
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest all rc=$?" | tee -a $O/pytest_gpu.log
  tail -8 $O/pytest_gpu.log
fi
