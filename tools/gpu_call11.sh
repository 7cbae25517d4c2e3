#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_variants.py 10000 2>&1 | tee gpurun_out/variants11.txt
for v in uncalled_b200/variants_pt/*.so; do
  timeout 200 python tools/gpu_phases.py g4m7 2368 "$v" > "gpurun_out/phases11_$(basename "$v" .so).txt" 2>&1; sed -n 3,26p "gpurun_out/phases11_$(basename "$v" .so).txt"
done
