#!/bin/bash
# One GPU call at the end of a round: parity tests, bench, phase breakdown, DRAM traffic of the
# full-size k2_map launch, launch list, ncu --set full of both kernels, layout experiment.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
if [ "$1" = "ref" ]; then timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; fi
if [ -f uncalled_b200/libunc_b200_pt.so ]; then timeout 300 python tools/gpu_phases.py g4m7 2368 > gpurun_out/phases.txt 2>&1; tail -12 gpurun_out/phases.txt; fi
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k2_map -s 1 -c 1 --csv --log-file gpurun_out/k2_traffic_full.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --reads 10000 > gpurun_out/k2_traffic_full.log 2>&1
tail -3 gpurun_out/k2_traffic_full.csv | cut -c1-60,380-
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_map -s 1 -c 1 -o gpurun_out/k2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --reads 1184 > gpurun_out/k2_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_events -s 1 -c 1 -o gpurun_out/k1_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --reads 10000 > gpurun_out/k1_full.log 2>&1
if ls uncalled_b200/variants/*.so > /dev/null 2>&1; then python tools/gpu_variants.py 10000 2>&1 | tee gpurun_out/variants_final.txt; fi
ls -la gpurun_out | head -30
