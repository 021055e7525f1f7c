#!/bin/bash
# build_variant.sh NPP NBUF UNIT PADW [extra nvcc flags] -> build/var/libtolcuda_nppN_nbufN_unitN_padN.so
# (record-buffer layout experiments; needs `make -C tol_b200/csrc` first: every object but fg_kernels.o is reused)
set -e
cd "$(dirname "$0")/../../tol_b200/csrc"
mkdir -p ../../build/var
nm=npp$1_nbuf$2_unit$3_pad$4
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false -Xcompiler -fPIC -DTOLCUDA_NPP=$1 -DTOLCUDA_NBUF=$2 -DTOLCUDA_UNIT=$3 -DTOLCUDA_PADW=$4 $5 -c fg_kernels.cu -o ../../build/var/fg_$nm.o 2>&1 | grep -v deprecated || true
others=$(ls ../../build/csrc/*.o | grep -v fg_kernels.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/var/libtolcuda_$nm.so ../../build/var/fg_$nm.o $others -Xcompiler -pthread 2>&1 | grep -v deprecated || true
echo built $nm "(run with TOLCUDA_LIB=build/var/libtolcuda_$nm.so)"
