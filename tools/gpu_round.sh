#!/bin/bash
# One gpurun call: GPU parity tests, the bench (both arms), the ncu launch list and one full capture of the match kernel.
# usage (here): gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi > $O/nvsmi.txt 2>&1
nproc > $O/nproc.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -3 $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json
timeout 600 python tools/perf_match.py 1000000 lane,staged > $O/perf_match.log 2>&1; cat $O/perf_match.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:match_kernel -s 3 -c 1 -f -o $O/match_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O
