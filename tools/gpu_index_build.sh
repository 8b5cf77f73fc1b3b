#!/bin/bash
# `krepp index` on the GPU: the new tests, then a library of G synthetic 3 Mbp genomes built by krepp_b200
# and by the unmodified reference from the same FASTA files -- wall times, the two libraries compared (offsets and encodings byte
# for byte, colours by expansion), and `dist` from either.   usage: gpurun -- 'bash tools/gpu_index_build.sh <tag> [genomes] [nopytest]'
TAG=${1:-idx}; G=${2:-64}; O=$PWD/gpurun_out/$TAG; mkdir -p $O
ROOT=$PWD; T=$(nproc)
if [ "$3" != "nopytest" ]; then
  ( time timeout 300 python -m pytest tests/test_gpu_index_build.py -x -q ) > $O/pytest_index_build.log 2>&1; echo "pytest rc=$?" >> $O/pytest_index_build.log
  tail -5 $O/pytest_index_build.log
fi
D=/tmp/idx_scale; rm -rf $D
tools/_build/synth_index --out $D --genomes $G --length 3000000 --fasta --reads 100000 --fastq-reads 100000 --seed 7 --threads $T > /dev/null 2>&1
cd $D
{
  echo "== $G genomes x 3 Mbp, -k 27 -w 35 -h 11, $T host threads"
  for rep in 1 2; do
    rm -rf gpu_index; TIMEFORMAT="krepp_b200 index run $rep: %R s wall"; time ( timeout 280 $ROOT/krepp_b200/_build/krepp_b200 --num-threads $T --verbose index -k 27 -w 35 -h 11 -o gpu_index -i input_map.tsv -t tree.nwk 2>&1 | grep -E "elapsed|k-mers:|stages|ERROR" )
  done
  TIMEFORMAT="reference krepp index: %R s wall"; time ( timeout 600 $ROOT/oracle/_ref/krepp --num-threads $T index -k 27 -w 35 -h 11 -o ref_index -i input_map.tsv -t tree.nwk 2>&1 | tr '\r' '\n' | grep -E "elapsed|ERROR" )
  python - <<PY
import sys, time
sys.path.insert(0, "$ROOT/tests")
from libraries import read_library, same_colours_vectorised
a, b, c = read_library("gpu_index"), read_library("ref_index"), read_library("index")
for f in ("metadata", "inc_bytes", "reflist", "tree"):
    print(f, "identical to the reference's:", a[f] == b[f])
print("k-mers:", a["nkmers"], b["nkmers"], "encoding column identical:", bool((a["enc"] == b["enc"]).all()))
print("rho identical:", a["rho"].tobytes() == b["rho"].tobytes())
print("colour ids: krepp_b200", a["nsubsets"], "reference", b["nsubsets"], "synth_index", c["nsubsets"])
t = time.time(); print("every k-mer's colour expands to the same references as the reference's:", same_colours_vectorised(a, b), f"({time.time() - t:.0f} s)")
print("... and as the generator's:", same_colours_vectorised(a, c))
PY
  for ix in gpu_index ref_index; do
    $ROOT/krepp_b200/_build/krepp_b200 --num-threads $T dist -i $ix -q reads.fq -o /tmp/d_$ix.tsv 2>&1 | grep -E "elapsed|ERROR"
    tail -n +3 /tmp/d_$ix.tsv | sort > /tmp/d_$ix.sorted
  done
  cmp /tmp/d_gpu_index.sorted /tmp/d_ref_index.sorted && echo "krepp_b200 dist: $(wc -l < /tmp/d_gpu_index.sorted) lines identical from the GPU-built and the reference-built library"
} 2>&1 | tee $O/index_scale.txt
