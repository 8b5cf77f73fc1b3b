#!/bin/bash
# `krepp_b200 index` at the size of the benchmark's configuration 3: 1,000 synthetic genomes x 3 Mbp on a random binary tree.
# The library is compared with the one tools/synth_index writes from the same genomes (offsets and encodings byte for byte,
# colours by expansion on a sample; rho differs by construction: HyperLogLog here as in the reference, exact in the generator).
# usage: gpurun -- 'bash tools/gpu_index_c3.sh <tag> [genomes]'
TAG=${1:-idxc3}; G=${2:-1000}; O=$PWD/gpurun_out/$TAG; mkdir -p $O
ROOT=$PWD; T=$(nproc); D=/tmp/idx_c3; rm -rf $D
{
  TIMEFORMAT="tools/synth_index (CPU generator, $T threads; also writes the FASTA files and reads): %R s wall"
  time ( tools/_build/synth_index --out $D --genomes $G --length 3000000 --fasta --reads 1000 --fastq-reads 1000 --seed 7 --threads $T > /dev/null 2>&1 )
  cd $D; du -sh genomes | sed 's/^/FASTA input: /'
  echo "== $G genomes x 3 Mbp, -k 27 -w 35 -h 11, $T host threads"
  for rep in 1 2; do
    rm -rf gpu_index; TIMEFORMAT="krepp_b200 index run $rep: %R s wall"; time ( timeout 280 $ROOT/krepp_b200/_build/krepp_b200 --num-threads $T --verbose index -k 27 -w 35 -h 11 -o gpu_index -i input_map.tsv -t tree.nwk 2>&1 | grep -E "elapsed|k-mers:|stages|ERROR" )
  done
  nvidia-smi --query-gpu=memory.used --format=csv,noheader | sed 's/^/HBM in use after the runs: /'
  timeout 200 python - <<PY
import sys, time
sys.path.insert(0, "$ROOT/tests")
from libraries import read_library, sampled_colour_check
a, c = read_library("gpu_index"), read_library("index")
print("metadata / offsets identical to the generator's:", a["metadata"] == c["metadata"], a["inc_bytes"] == c["inc_bytes"])
print("k-mers:", a["nkmers"], c["nkmers"], "encoding column identical:", bool((a["enc"] == c["enc"]).all()))
print("colour ids: krepp_b200", a["nsubsets"], "generator", c["nsubsets"])
t = time.time(); n, bad = sampled_colour_check(a, c); print(f"colours of {n} sampled k-mers expanded in both libraries: {bad} differ ({time.time() - t:.0f} s)")
PY
} 2>&1 | tee $O/index_c3.txt
