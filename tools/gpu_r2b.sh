#!/bin/bash
# quick iteration + the drop-in executable on config 3 with stage timers.  usage: gpurun -- 'bash tools/gpu_r2b.sh <tag> [reads]'
TAG=${1:-q}; N=${2:-1000000}
O=gpurun_out/$TAG; mkdir -p $O
TESTS="tests/test_gpu_shim.py tests/test_gpu_sorted.py tests/test_gpu_c3.py" bash tools/gpu_quick2.sh $TAG
D=$(python -c "import sys; sys.path.insert(0,'tools'); import workload as W; print(W.ensure_c3($N, fastq_reads=$N)[0])")
T=$(nproc)
for mode in dist place; do
  for rep in 1 2; do
    ( TIMEFORMAT="wall %R s"; time krepp_b200/_build/krepp_b200 --verbose --num-threads $T $mode -i $D/index -q $D/reads.fq -o /tmp/gpu_$mode.out ) 2>&1 | grep -E "stages|elapsed|wall" | sed "s/^/$mode run $rep: /"
  done
done | tee $O/cli_stages.txt
( KREPP_READER_DEBUG=1 krepp_b200/_build/krepp_b200 --num-threads $T dist -i $D/index -q $D/reads.fq -o /tmp/gpu_dist.out ) 2>&1 | grep "\[reader\]" | tee -a $O/cli_stages.txt
