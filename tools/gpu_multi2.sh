#!/bin/bash
# Multi-GPU round: the default bench line at N GPUs as the driver launches it (dist + place + mode_b objects), the reference
# arm, and the multi-GPU tests.  usage: gpurun --gpus N --timeout 2400 -- 'bash tools/gpu_multi2.sh <tag> N [steps]'
TAG=${1:-m}; N=${2:-2}; STEPS=${3:-5}
O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=index,name,memory.used --format=csv > $O/nvsmi.txt 2>&1
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps $STEPS --warmup 3 ) > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"
tail -4 $O/bench_n$N.err
python - $O/bench_n$N.json <<'PY'
import json, sys
line = [l for l in open(sys.argv[1]) if l.startswith("{")][-1]
d = json.loads(line)
def show(tag, x):
    if not x: return
    e = x.get("e2e") or {}
    print("%-8s value %8.2f M reads/s   e2e %8.2f M   ms/step %7.1f   d2h/read %s" % (tag, x["value"] / 1e6, e.get("value", 0) / 1e6, x["ms_per_step"], e.get("d2h_bytes_per_read")))
show("dist", d); show("place", d.get("place")); show("mode_b", d.get("mode_b"))
mb = d.get("mode_b")
if mb:
    print("mode_b budget:", mb["config"]["budget"]); print("mode_b exchange:", mb["exchange"]); print("mode_b stages:", mb["roofline"]["last_batch_stages_ms_rank0"])
PY
( time timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_parity.py -m gpu -q -x -k "shard or two_gpus" ) > $O/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_multi.log
