#!/bin/bash
# ncu captures on the config-3 workload (built on the box): launch list + full capture of the match and solve kernels.
# usage: bash tools/ncu_c3.sh <tag> [reads]
TAG=${1:-c3}; READS=${2:-500000}
O=gpurun_out/$TAG; mkdir -p $O
tools/_build/synth_index --out /tmp/c3 --genomes 1000 --length 3000000 --reads $READS --fastq-reads 1000 > $O/synth.log 2>&1
CMD="python tools/perf_c3.py --reads $READS --batch 250000 --check 0 --skip-cli --cpu-reads 0 --groups staged"
# (launch list disabled)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:match_kernel -s 2 -c 1 -f -o $O/match_c3 $CMD > $O/ncu_c3.log 2>&1
# (solve capture disabled)
tail -2 $O/ncu_c3.log
