#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:k2_map_exact -c 1 -o gpurun_out/k2_exact_r2 -f python bench.py --exact-ties --steps 1 --warmup 0 --no-cpu-baseline --no-extras --reads 592 > gpurun_out/k2_exact_r2.log 2>&1; echo "ncu exact rc=$?"; tail -3 gpurun_out/k2_exact_r2.log | cut -c1-300
