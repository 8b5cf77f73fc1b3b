#!/bin/bash
# Quick iteration on the placement chain: parity (config 3, golden, segments), then the place bench without the CPU baseline.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_place_quick.sh <tag>'
TAG=${1:-pq}; O=gpurun_out/$TAG; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_c3.py tests/test_gpu_parity.py tests/test_gpu_segments.py -m gpu -q -k "not cli" ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --mode place --reads 2000000 --batch 500000 --no-cpu-baseline > $O/bench_place.json 2>/dev/null
python - $O/bench_place.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("place value %.2f M e2e %.2f M ms/step %.1f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"]))
print(d["roofline"]["stages_ms_per_step"])
PY
