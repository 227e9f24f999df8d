#!/bin/sh
# Tuning build: libcgmres_b200_quick.so = the regular objects, except that the two third-generation translation
# units are recompiled with only the msd kernels instantiated (-DCG_ONLY_MSD, ~40 s instead of ~2 min) plus any extra
# flags given on the command line.  Load it with CGMRES_B200_LIB=$PWD/cgmres_cpp_b200/libcgmres_b200_quick.so.
set -e
cd "$(dirname "$0")/../cgmres_cpp_b200/csrc"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="$ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -I../../include -I. -DCG_ONLY_MSD $*"
mkdir -p ../_build_quick
make -s ../_build/capi.o ../_build/layout.o ../_build/peak.o > /dev/null  # host side: always current
rm -f ../_build_quick/*.o ../libcgmres_b200_quick.so
nvcc $COMMON -Xptxas -v -c pipe2_fast_kernels.cu -o ../_build_quick/pipe2_fast_kernels.o 2> ../_build_quick/pipe2_fast.log &
nvcc $COMMON -fmad=false -Xptxas -v -c pipe2_exact_kernels.cu -o ../_build_quick/pipe2_exact_kernels.o 2> ../_build_quick/pipe2_exact.log || true
wait
for f in fast exact; do
  if [ ! -f ../_build_quick/pipe2_${f}_kernels.o ]; then grep -m 5 -A3 "error" ../_build_quick/pipe2_$f.log; echo "BUILD FAILED ($f)"; exit 1; fi
done
grep -h -E "registers|spill" ../_build_quick/pipe2_*.log | sort | uniq -c
nvcc $ARCH -shared -o ../libcgmres_b200_quick.so ../_build/capi.o ../_build/layout.o ../_build/peak.o ../_build/exact_kernels.o \
  ../_build/onchip_exact_kernels.o ../_build/fast_kernels.o ../_build_quick/pipe2_fast_kernels.o ../_build_quick/pipe2_exact_kernels.o -lgomp
