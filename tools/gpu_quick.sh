#!/bin/bash
# Parity suite + default bench (c3, no reference arm) + place side line.  usage: gpurun --timeout 1800 -- 'bash tools/gpu_quick.sh <tag> [memo-off]'
TAG=${1:-q2}; O=gpurun_out/$TAG; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -3 $O/bench.err
timeout 600 python bench.py --mode place --reads 2000000 --batch 500000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/bench_place.json 2> $O/bench_place.err; echo "bench place rc=$?"; cat $O/bench_place.json; tail -3 $O/bench_place.err
if [ "$2" == "memo-off" ]; then KREPP_MEMO=0 timeout 900 python bench.py --no-cpu-baseline --no-e2e --steps 2 > $O/bench_nomemo.json 2> $O/bench_nomemo.err; echo "bench nomemo rc=$?"; cat $O/bench_nomemo.json; fi
