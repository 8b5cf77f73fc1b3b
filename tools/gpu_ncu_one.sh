#!/bin/bash
# One full ncu capture of selected kernels on the config-3 workload (one 250k-read batch after the warm-up batches).
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_ncu_one.sh <tag> <kernel regex> [launches per batch matching the regex] [extra bench.py arguments]'
TAG=${1:-ncu1}; RE=${2:-resolve_kernel}; PER=${3:-1}; EXTRA=${4:-}; O=gpurun_out/$TAG; mkdir -p $O
CMD="python bench.py --reads 1000000 --batch 250000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $EXTRA"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $((4 * PER)) -c $PER -f -o $O/one $CMD > $O/ncu.log 2>&1; echo "ncu rc=$?"
tail -3 $O/ncu.log
