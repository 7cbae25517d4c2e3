#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none -k regex:k_dtw -s 1 -c 1 -o gpurun_out/k_dtw_r2 -f python tools/bench_dtw.py --problems 128 --cpu-problems 0 > gpurun_out/k_dtw_r2.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/k_dtw_r2.log | cut -c1-300
