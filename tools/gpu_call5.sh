#!/bin/bash
# Round 2, fifth GPU call: fused sort/dedup phase (look-back) on hardware -- parity suite, shapes, phases, bench, ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log
tail -5 gpurun_out/pytest_gpu5.log
python tools/gpu_variants.py 10000 2>&1 | tee gpurun_out/variants5.txt
for v in uncalled_b200/variants_pt/*.so; do
  timeout 200 python tools/gpu_phases.py g4m7 2368 "$v" > "gpurun_out/phases5_$(basename "$v" .so).txt" 2>&1; tail -14 "gpurun_out/phases5_$(basename "$v" .so).txt"
done
timeout 900 python bench.py > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench5.json; tail -5 gpurun_out/bench5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_map -s 1 -c 1 -o gpurun_out/k2v24_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --reads 1184 > gpurun_out/k2v24_full.log 2>&1
ls -la gpurun_out | tail -6
