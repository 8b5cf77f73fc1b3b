#!/bin/bash
# Quick GPU check of a kernel change: parity tests, then the config-3 bench line without the CPU legs.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_quick.sh <tag> [ncu]'
TAG=${1:-q}; O=gpurun_out/$TAG; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -3 $O/bench.err
if [ "$2" == "ncu" ]; then
  CMD="python bench.py --reads 1000000 --batch 250000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:match_kernel -s 5 -c 1 -f -o $O/match_full $CMD > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
