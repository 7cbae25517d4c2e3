#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_variants.py 10000 2>&1 | tee gpurun_out/variants9.txt
for v in uncalled_b200/variants_pt/*.so; do
  timeout 200 python tools/gpu_phases.py g4m7 2368 "$v" > "gpurun_out/phases9_$(basename "$v" .so).txt" 2>&1; tail -15 "gpurun_out/phases9_$(basename "$v" .so).txt"
done
