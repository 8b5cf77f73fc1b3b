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
for v in "KREPP_OUT_DIRECT=0 -o /tmp/gpu_dist_pw.out" "KREPP_OUT_DIRECT=1 -o /dev/null"; do
  set -- $v
  ( TIMEFORMAT="wall %R s"; time env $1 krepp_b200/_build/krepp_b200 --verbose --num-threads $T dist -i $D/index -q $D/reads.fq $2 $3 ) 2>&1 | grep -E "stages|elapsed|wall" | sed "s|^|dist ($v): |"
done | tee -a $O/cli_stages.txt
cmp <(tail -n +2 /tmp/gpu_dist.out) <(tail -n +2 /tmp/gpu_dist_pw.out) && echo "writer-thread output == pwrite output" | tee -a $O/cli_stages.txt
( KREPP_READER_DEBUG=1 krepp_b200/_build/krepp_b200 --num-threads $T dist -i $D/index -q $D/reads.fq -o /tmp/gpu_dist.out ) 2>&1 | grep "\[reader\]" | tee -a $O/cli_stages.txt
