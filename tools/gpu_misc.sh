#!/bin/bash
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_misc.sh <tag>'   (CLI drop-in check on config 3, place parity + bench)
TAG=${1:-misc}; O=gpurun_out/$TAG; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sorted.py -x -q -k "place or golden" ) > $O/pytest_place.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_place.log
timeout 900 bash tools/gpu_cli_c3.sh $O 200000
timeout 600 python bench.py --mode place --reads 2000000 --batch 500000 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_place.json 2> $O/bench_place.err; echo "bench place rc=$?"; cut -c1-300 $O/bench_place.json
