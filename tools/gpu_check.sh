#!/bin/bash
# Lean GPU check: parity tests + the default bench line (no reference arm, no ncu).
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh <tag>'
TAG=${1:-chk}; O=gpurun_out/$TAG; mkdir -p $O
( nproc; free -g; df -h /tmp /dev/shm; nvidia-smi -L ) > $O/box.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -3 $O/bench.err
