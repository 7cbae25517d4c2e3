#!/bin/bash
mkdir -p gpurun_out
for v in uncalled_b200/variants_pt/*.so; do
  timeout 200 python tools/gpu_phases.py g4m7 2368 "$v" > "gpurun_out/phases10_$(basename "$v" .so).txt" 2>&1; tail -50 "gpurun_out/phases10_$(basename "$v" .so).txt"
done
