#!/bin/bash
# Quick GPU check of a kernel change: the GPU parity tests, then config-3 timings with the per-stage breakdown.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_perf.sh <tag> [reads] [extra perf_c3 flags]'
TAG=${1:-p}; READS=${2:-2000000}; shift 2
O=gpurun_out/$TAG; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
timeout 1200 python tools/perf_c3.py --reads $READS --batch 1000000 --check 300 --skip-cli --cpu-reads 0 --pipelines sorted --groups staged "$@" > $O/perf_c3.log 2>&1; echo "perf_c3 rc=$?"
tail -14 $O/perf_c3.log
