#!/bin/bash
# Long queries (contigs / long reads) on the 1,000-genome configuration through the drop-in executable: cut into segments (default),
# uncut (KREPP_SEGMENT_WINDOWS=0: one warp per read through lookup and resolve), and the reference CLI on all host threads; the three
# outputs are compared.  usage: gpurun -- 'bash tools/gpu_long.sh <tag>'
TAG=${1:-long}; O=gpurun_out/$TAG; mkdir -p $O
T=$(nproc)
EXE=krepp_b200/_build/krepp_b200
{
for spec in "30000 3000" "1000000 64"; do
  set -- $spec; L=$1; N=$2; D=/tmp/c3long_$L
  [ -f $D/workload.json ] || tools/_build/synth_index --out $D --genomes 1000 --length 3000000 --seed 7 --reads $N --read-len $L --fastq-reads $N > /dev/null
  echo "== $N reads of $L bp ($(( L * N / 1000000 )) Mbp), $T host threads"
  for rep in 1 2; do
    ( TIMEFORMAT="wall %R s"; time $EXE --verbose --num-threads $T dist -i $D/index -q $D/reads.fq -o /tmp/long_cut.tsv ) 2>&1 | grep -E "stages|elapsed|wall" | sed "s/^/cut run $rep: /"
  done
  ( TIMEFORMAT="wall %R s"; time KREPP_SEGMENT_WINDOWS=0 timeout 600 $EXE --verbose --num-threads $T dist -i $D/index -q $D/reads.fq -o /tmp/long_whole.tsv ) 2>&1 | grep -E "stages|elapsed|wall" | sed "s/^/whole reads: /"
  ( TIMEFORMAT="wall %R s"; time oracle/_ref/krepp --num-threads $T dist -i $D/index -q $D/reads.fq -o /tmp/long_ref.tsv ) 2>&1 | grep -E "elapsed|wall" | sed "s/^/reference: /"
  for f in cut whole ref; do tail -n +3 /tmp/long_$f.tsv | sort > /tmp/long_$f.sorted; done
  echo "rows: $(wc -l < /tmp/long_cut.sorted) cut, $(wc -l < /tmp/long_whole.sorted) whole, $(wc -l < /tmp/long_ref.sorted) reference"
  cmp /tmp/long_cut.sorted /tmp/long_whole.sorted && echo "cut == whole"
  cmp /tmp/long_cut.sorted /tmp/long_ref.sorted && echo "cut == reference" || diff /tmp/long_cut.sorted /tmp/long_ref.sorted | head -6
done
} 2>&1 | tee $O/long_reads.txt
