#!/bin/bash
# Last GPU call of round 2: the whole -m gpu suite with the index builder in, smoke, compute-sanitizer memcheck on the builder's
# kernels, and ncu on one 64-genome `krepp_b200 index` (launch list + --set full of the builder kernels and one minimizer launch).
# usage: gpurun --timeout 660 -- 'bash tools/gpu_r12z.sh <tag>'
TAG=${1:-r12z}; O=$PWD/gpurun_out/$TAG; mkdir -p $O; ROOT=$PWD; T=$(nproc)
( time timeout 420 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
SAN=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
timeout 150 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -x -q -m gpu tests/test_gpu_index_build.py::test_builder_through_the_c_abi tests/test_gpu_index_build.py::test_library_of_draft_assemblies > $O/memcheck.log 2>&1; echo "memcheck rc=$?"
grep -n "ERROR SUMMARY\|passed\|failed" $O/memcheck.log | head -5
D=/tmp/idx_ncu; rm -rf $D
tools/_build/synth_index --out $D --genomes 64 --length 3000000 --fasta --reads 1000 --fastq-reads 1000 --seed 7 --threads $T > /dev/null 2>&1
cd $D
CMD="$ROOT/krepp_b200/_build/krepp_b200 --num-threads $T index -k 27 -w 35 -h 11 -o gpu_index -i input_map.tsv -t tree.nwk"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
rm -rf gpu_index
timeout 120 ncu --set full --clock-control none --import-source on -k "regex:head_kernel|run_start_kernel|set_hash_kernel|group_head_kernel|set_rep_kernel|set_assign_kernel|set_gather_kernel" -c 8 -f -o $O/builder_full $CMD > $O/ncu_builder.log 2>&1; echo "ncu builder rc=$?"
rm -rf gpu_index
timeout 100 ncu --set full --clock-control none --import-source on -k "regex:minimizer_kernel" -s 5 -c 1 -f -o $O/minimizer_full $CMD > $O/ncu_minimizer.log 2>&1; echo "ncu minimizer rc=$?"
ls -la $O
