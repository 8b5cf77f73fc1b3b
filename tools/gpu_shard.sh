#!/bin/bash
# Mode B + place check on one GPU: the shard parity tests (logical ranks on cuda:0), the place side line (configs[3]) and the
# sharded bench arm with a single shard (plumbing check).  usage: gpurun --timeout 1500 -- 'bash tools/gpu_shard.sh <tag>'
TAG=${1:-shard}; O=gpurun_out/$TAG; mkdir -p $O
( nproc; free -g; nvidia-smi -L ) > $O/box.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_shard.py -x -q ) > $O/pytest_shard.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_shard.log
tail -15 $O/pytest_shard.log
timeout 600 python bench.py --mode place --reads 2000000 --batch 500000 --steps 2 --warmup 1 --cpu-sample 50000 > $O/bench_place.json 2> $O/bench_place.err; echo "bench place rc=$?"; cat $O/bench_place.json; tail -3 $O/bench_place.err
timeout 600 python bench.py --workload c5 --reads 2000000 --steps 2 --warmup 1 > $O/bench_c5_n1.json 2> $O/bench_c5_n1.err; echo "bench c5 rc=$?"; cat $O/bench_c5_n1.json; tail -3 $O/bench_c5_n1.err
