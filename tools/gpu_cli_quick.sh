#!/bin/bash
# The drop-in executable alone on 1 M config-3 reads: output written by the formatter threads (pwrite, default), by one writer
# thread (KREPP_OUT_DIRECT=0), and to /dev/null; dist and place.  usage: gpurun -- 'bash tools/gpu_cli_quick.sh <tag>'
TAG=${1:-cli}; N=${2:-1000000}; O=gpurun_out/$TAG; mkdir -p $O
D=$(python -c "import sys; sys.path.insert(0,'tools'); import workload as W; print(W.ensure_c3($N, fastq_reads=$N)[0])" 2>/dev/null | tail -1)
T=$(nproc)
for mode in dist place; do
  for v in "KREPP_OUT_DIRECT=1 /tmp/cli_a.out" "KREPP_OUT_DIRECT=0 /tmp/cli_b.out" "KREPP_OUT_DIRECT=1 /dev/null"; do
    set -- $v
    for rep in 1 2; do
      ( env $1 krepp_b200/_build/krepp_b200 --verbose --num-threads $T $mode -i $D/index -q $D/reads.fq -o $2 ) 2>&1 | grep -E "stages|elapsed" | sed "s|^|$mode $v run $rep: |"
    done
  done
  # (the jplace footer quotes the invocation, which names the output file: left out of the comparison)
  cmp <(tail -n +2 /tmp/cli_a.out | grep -v '"invocation"') <(tail -n +2 /tmp/cli_b.out | grep -v '"invocation"') && echo "$mode: writer-thread output == pwrite output"
done 2>&1 | tee $O/cli_writers.txt
