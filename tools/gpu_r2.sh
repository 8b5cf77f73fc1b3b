#!/bin/bash
# Round-2 GPU call: parity tests (all, no -x so that every failure is seen), smoke, the default bench line (dist + place objects)
# and the reference arm.  usage: gpurun --timeout 2400 -- 'bash tools/gpu_r2.sh <tag> [pytest-args]'
TAG=${1:-r07a}; shift
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi > $O/nvsmi.txt 2>&1
( nproc; free -g; df -h /tmp /dev/shm ) > $O/box.txt 2>&1; cat $O/box.txt
( time timeout 1800 python -m pytest tests -m gpu -q "$@" ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -5 $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"; cut -c1-400 $O/bench_ref.json; tail -3 $O/bench_ref.err
ls -la $O
