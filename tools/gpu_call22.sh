#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_variants.py 10000 2>&1 | tee gpurun_out/variants22.txt
for v in uncalled_b200/variants_pt/*.so; do
  n=$(basename "$v" .so)
  timeout 200 python tools/gpu_phases.py g4m7 2368 "$v" > "gpurun_out/phases22_$n.txt" 2>&1; echo "== $v"; sed -n 3,32p "gpurun_out/phases22_$n.txt"
done
