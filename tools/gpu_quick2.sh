#!/bin/bash
# Quick kernel iteration: the bucket-sorted chain's parity tests, then the dist bench without the place / cpu-baseline legs.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_quick2.sh <tag> [extra bench args]'
TAG=${1:-q}; shift
O=gpurun_out/$TAG
mkdir -p $O
( time timeout 900 python -m pytest ${TESTS:-tests/test_gpu_sorted.py tests/test_gpu_c3.py} -m gpu -q -x -k "not cli" ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
timeout 600 python bench.py --no-place --no-cpu-baseline "$@" > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - $O/bench.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.2f M reads/s  e2e %.2f M  ms/step %.1f  whole_step_frac %.3f" % (d["value"] / 1e6, (d["e2e"] or {}).get("value", 0) / 1e6, d["ms_per_step"], d["roofline"]["whole_step_frac"]))
for k, v in d["roofline"]["stages_ms_per_step"].items():
    print("  %-40s %8.3f" % (k, v))
PY
