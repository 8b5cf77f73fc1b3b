#!/bin/bash
# Two or more GPUs: the command-line tests (incl. --num-gpus and --shard-index over real peer copies), then the sharded command line
# against the replicated one on the config-3 workload (same bytes expected).  usage: gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_cli_shard.sh <tag> <N>'
TAG=${1:-clishard}; N=${2:-2}; O=gpurun_out/$TAG; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k cli ) > $O/pytest_cli.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_cli.log
D=$(python -c "import sys; sys.path.insert(0,'tools'); import workload as W; print(W.ensure_c3(400000, fastq_reads=400000)[0])")
T=$(nproc)
( TIMEFORMAT="wall %R s"; time krepp_b200/_build/krepp_b200 --num-threads $T dist -i $D/index -q $D/reads.fq -o /tmp/one.tsv ) > $O/one.log 2>&1
( TIMEFORMAT="wall %R s"; time krepp_b200/_build/krepp_b200 --num-threads $T dist -i $D/index -q $D/reads.fq -o /tmp/rep.tsv --num-gpus $N ) > $O/rep.log 2>&1
( TIMEFORMAT="wall %R s"; time krepp_b200/_build/krepp_b200 --num-threads $T dist -i $D/index -q $D/reads.fq -o /tmp/shard.tsv --num-gpus $N --shard-index --batch-reads 100000 ) > $O/shard.log 2>&1
( echo "# krepp_b200 dist on 400,000 reads of the config-3 workload, $N GPUs"; for f in one rep shard; do echo "$f: $(grep -h 'elapsed\|wall' $O/$f.log | tr '\n' ' ')"; done
  tail -n +2 /tmp/one.tsv > /tmp/one.body
  tail -n +2 /tmp/rep.tsv | cmp /tmp/one.body - && echo "replicated over $N GPUs: output identical to one GPU (first line aside: it carries the invocation)"
  tail -n +2 /tmp/shard.tsv | cmp /tmp/one.body - && echo "index sharded over $N GPUs: output identical to one GPU"; wc -l /tmp/one.body ) | tee $O/cli_shard.txt
