#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the bench (config 3 default, both arms; toy = configs[1] as a side line), the
# ncu launch list and one full capture of the match kernel on the config-3 workload.
# usage (here): gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag] [skip-ncu]'
TAG=${1:-r01}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi > $O/nvsmi.txt 2>&1
( nproc; free -g; df -h /tmp /dev/shm ) > $O/box.txt 2>&1; cat $O/box.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?"; cat $O/bench_ref.json; tail -3 $O/bench_ref.err
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -3 $O/bench.err
timeout 600 python bench.py --workload toy --no-cpu-baseline > $O/bench_toy.json 2> $O/bench_toy.err; echo "bench toy rc=$?"; cat $O/bench_toy.json; tail -3 $O/bench_toy.err
timeout 600 python bench.py --mode place --reads 2000000 --batch 500000 --cpu-sample 50000 > $O/bench_place.json 2> $O/bench_place.err; echo "bench place rc=$?"; cut -c1-200 $O/bench_place.json
timeout 900 bash tools/gpu_cli_c3.sh $O 1000000 > $O/cli.log 2>&1; cat $O/cli_c3.txt
if [ "$2" != "skip-ncu" ]; then
  CMD="python bench.py --reads 2000000 --batch 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"  # launches of 1M reads, as the bench line
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  # one batch's worth of the heavy kernels (8 matching launches per batch; the 2 warm-up batches are skipped)
  timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:lookup_kernel|join_kernel|hit_scatter_kernel|resolve_kernel|gate_kernel|solve_kernel|alias_kernel" -s 16 -c 8 -f -o $O/chain_full $CMD > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
ls -la $O
