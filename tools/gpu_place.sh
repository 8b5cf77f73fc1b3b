#!/bin/bash
# Full GPU parity suite + the place side line (configs[3]).  usage: gpurun --timeout 1800 -- 'bash tools/gpu_place.sh <tag>'
TAG=${1:-place}; O=gpurun_out/$TAG; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
timeout 600 python bench.py --mode place --reads 2000000 --batch 500000 --steps 3 --warmup 2 --cpu-sample 50000 > $O/bench_place.json 2> $O/bench_place.err; echo "bench place rc=$?"; cat $O/bench_place.json; tail -3 $O/bench_place.err
timeout 600 python bench.py --workload c5 --reads 2000000 --steps 2 --warmup 1 --no-e2e > $O/bench_c5_n1.json 2> $O/bench_c5_n1.err; echo "bench c5 rc=$?"; cat $O/bench_c5_n1.json; tail -3 $O/bench_c5_n1.err
