#!/bin/bash
# Round-2 candidate builds of the mapper kernel for tools/gpu_variants.py: the first structure (K2_V1) as the
# reference point, the second structure (unc_k2v2.cuh) at 2 / 3 / 4 CTAs x 8 warps and 2 CTAs x 12 warps per SM.
set -e
cd "$(dirname "$0")/../uncalled_b200"
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC --shared -diag-suppress 550"
rm -rf variants variants_pt; mkdir -p variants variants_pt
SRC="csrc/unc_abi.cu csrc/unc_index_build.cpp csrc/unc_fast5.cpp -lz"
build() { nvcc $F "${@:2}" -o "variants/$1.so" $SRC; }
build a_v1 -DK2_V1 -DK2_MIN_CTAS=2 &
build v2_c2 -DK2_MIN_CTAS=2 &
build v2_c3 -DK2_MIN_CTAS=3 &
build v2_c4 -DK2_MIN_CTAS=4 &
wait
build v2_w12c2 -DK2_WARPS=12 -DK2_MAXSEG=16u -DK2_MIN_CTAS=2 &
build v2_w16c1 -DK2_WARPS=16 -DK2_MAXSEG=16u -DK2_MIN_CTAS=1 &
nvcc $F -DUNC_PHASE_TIMING -DK2_MIN_CTAS=3 -o variants_pt/v2_c3.so $SRC &
wait
ls -la variants variants_pt
