#!/bin/bash
# N GPUs of one box: [cli] the config-3 drop-in check, [tests] the shard parity tests (incl. the NCCL two-GPU test), then the
# sharded arm (mode B) and the replicated arm (mode A) under torchrun.
# usage: gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi.sh <tag> <N> [cli] [tests]'
TAG=${1:-shard2}; N=${2:-2}; O=gpurun_out/$TAG; mkdir -p $O
( nproc; free -g; nvidia-smi -L; nvidia-smi topo -m ) > $O/box.txt 2>&1
if [[ " $* " == *" tests "* ]]; then
  ( time timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_parity.py::test_cli_two_gpus_same_bytes -x -q ) > $O/pytest_shard.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_shard.log; tail -4 $O/pytest_shard.log
fi
if [[ " $* " == *" cli "* ]]; then timeout 900 bash tools/gpu_cli_c3.sh $O 200000; fi
python -c "import sys; sys.path.insert(0,'tools'); import workload as W; W.ensure_c3(4000000*$N)" > $O/workload.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --workload c5 --gpus $N --steps 3 --warmup 2 > $O/bench_c5_n$N.json 2> $O/bench_c5_n$N.err; echo "bench c5 N=$N rc=$?"; cat $O/bench_c5_n$N.json | cut -c1-400; grep -v Warn $O/bench_c5_n$N.err | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --reads 4000000 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_c3_n$N.json 2> $O/bench_c3_n$N.err; echo "bench c3 N=$N rc=$?"; cat $O/bench_c3_n$N.json | cut -c1-400; grep -v Warn $O/bench_c3_n$N.err | tail -3
