#!/bin/bash
# Builds the candidate configurations of the mapper kernel into uncalled_b200/variants/ for
# tools/gpu_variants.py (one GPU call times them all on the bench workload and checks that their PAF
# records are identical).  ptxas figures of round 1 (sm_100a, CUDA 12.9), k2_map:
#   default                      124 regs, 0 B spilled   -> 2 CTAs x 8 warps / SM  (shipped)
#   default   @ 3 CTAs (80 regs) 252 B spilled           -> measured 15-20 % slower
#   K2_LEAN_B @ 3 CTAs (80 regs)   0 B spilled
#   K2_LEAN_B @ 4 CTAs (64 regs)  32 B spilled           (also 2 CTAs x 16 warps)
#   K2_TRK_INLINE (no tracker warp) 128 regs, 16 B spilled, all in per-read set-up code
#   K2_TRK_INLINE + LEAN_B + PAR_E + SCAN2 @ 3 CTAs: 20 B spilled, all in per-read set-up / record code
set -e
cd "$(dirname "$0")/../uncalled_b200"
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC --shared -diag-suppress 550"
mkdir -p variants
SRC="csrc/unc_abi.cu csrc/unc_index_build.cpp csrc/unc_fast5.cpp -lz"
build() { nvcc $F "${@:2}" -o "variants/$1.so" $SRC; }
build base &
build lean_c2 -DK2_LEAN_B -DK2_MIN_CTAS=2 &
build lean_c3 -DK2_LEAN_B -DK2_MIN_CTAS=3 &
build lean_c4 -DK2_LEAN_B -DK2_MIN_CTAS=4 &
wait
build lean_w16c2 -DK2_LEAN_B -DK2_WARPS=16 -DK2_MAXSEG=16u -DK2_MIN_CTAS=2 &
build lean_w12c2 -DK2_LEAN_B -DK2_WARPS=12 -DK2_MAXSEG=16u -DK2_MIN_CTAS=2 &
build lean_pare_c3 -DK2_LEAN_B -DK2_PAR_E -DK2_MIN_CTAS=3 &
build lean_pare_w16c2 -DK2_LEAN_B -DK2_PAR_E -DK2_WARPS=16 -DK2_MAXSEG=16u -DK2_MIN_CTAS=2 &
wait
build scan2 -DK2_SCAN2 &
build trk -DK2_TRK_INLINE &
build trk_scan2 -DK2_TRK_INLINE -DK2_SCAN2 &
build pf2 -DK2_PF2 &
build bmatch -DK2_BMATCH &
build dfuse -DK2_DFUSE &
build trk_pf2 -DK2_TRK_INLINE -DK2_PF2 &
build lean_pare_scan2_c3 -DK2_LEAN_B -DK2_PAR_E -DK2_SCAN2 -DK2_MIN_CTAS=3 &
wait
# no tracker warp (every warp works) on top of the lean loop: 3 x 8, 4 x 8 and 2 x 16 working warps per SM
ALL="-DK2_TRK_INLINE -DK2_LEAN_B -DK2_PAR_E -DK2_SCAN2"
build trk_lean_c3 -DK2_TRK_INLINE -DK2_LEAN_B -DK2_MIN_CTAS=3 &
build trk_all_c3 $ALL -DK2_MIN_CTAS=3 &
build trk_all_c4 $ALL -DK2_MIN_CTAS=4 &
build trk_all_w16c2 $ALL -DK2_WARPS=16 -DK2_MAXSEG=16u -DK2_MIN_CTAS=2 &
wait
build trk_all_pf2_c3 $ALL -DK2_PF2 -DK2_MIN_CTAS=3 &
build trk_all_pf2_dfuse_c3 $ALL -DK2_PF2 -DK2_DFUSE -DK2_MIN_CTAS=3 &
build trk_all_pf2_dfuse_bmatch_c3 $ALL -DK2_PF2 -DK2_DFUSE -DK2_BMATCH -DK2_MIN_CTAS=3 &
build trk_all_pf2_w16c2 $ALL -DK2_PF2 -DK2_WARPS=16 -DK2_MAXSEG=16u -DK2_MIN_CTAS=2 &
wait
# phase-timing builds (tools/gpu_phases.py <genome> <reads> <lib>)
mkdir -p variants_pt
nvcc $F -DUNC_PHASE_TIMING -o libunc_b200_pt.so $SRC &      # the shipped configuration with phase marks
nvcc $F -DUNC_PHASE_TIMING -DK2_LEAN_B -DK2_PAR_E -DK2_WARPS=16 -DK2_MAXSEG=16u -DK2_MIN_CTAS=2 -o variants_pt/lean_pare_w16c2.so $SRC &
nvcc $F -DUNC_PHASE_TIMING -DK2_LEAN_B -DK2_MIN_CTAS=3 -o variants_pt/lean_c3.so $SRC &
nvcc $F -DUNC_PHASE_TIMING $ALL -DK2_MIN_CTAS=3 -o variants_pt/trk_all_c3.so $SRC &
nvcc $F -DUNC_PHASE_TIMING -DK2_TRK_INLINE -o variants_pt/trk.so $SRC &
wait
ls -la variants variants_pt
