#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer over the HOST side of the library (index loader, reader / writers, library
# writer, and the host code of api.cu / builder.cu / minimizer.cu) with the CPU test suite -- no GPU needed.  An instrumented copy
# of libkrepp_b200.so is linked in /tmp/asan (device code and the three big kernel files as built) and the tests load it instead
# of krepp_b200/_build/libkrepp_b200.so.  usage: bash tools/asan_cpu.sh [pytest args]   (after make -C krepp_b200/csrc)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd); A=/tmp/asan; mkdir -p $A; cd $ROOT/krepp_b200/csrc
SAN="-fsanitize=address -fsanitize=undefined -fno-omit-frame-pointer"
for f in library_writer index_image host_io; do /usr/bin/g++ -std=c++17 -O1 -g -fPIC -Wall $SAN -I/usr/local/cuda/include -c $f.cpp -o $A/$f.o; done
for f in api builder minimizer; do /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O1 -g -Xcompiler -fPIC,${SAN// /,} -ccbin /usr/bin/g++ -c $f.cu -o $A/$f.o; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $A/libkrepp_b200.so ../_build/match.o ../_build/sorted.o ../_build/solve.o $A/minimizer.o $A/builder.o \
  $A/library_writer.o $A/api.o $A/index_image.o $A/host_io.o -ccbin /usr/bin/g++ -lz -Xlinker -lasan -Xlinker -lubsan
cat > $A/run.py <<PY
import sys
sys.path[:0] = ["$ROOT", "$ROOT/tests", "$ROOT/tools"]
import krepp_b200.capi as c
c._LIB = "$A/libkrepp_b200.so"
import pytest
sys.exit(pytest.main(sys.argv[1:]))
PY
cd $ROOT
LD_PRELOAD=$(/usr/bin/g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1:protect_shadow_gap=0 UBSAN_OPTIONS=print_stacktrace=1 \
  python $A/run.py tests -m "not gpu" -x -s -q -k "not gloo" "$@"
