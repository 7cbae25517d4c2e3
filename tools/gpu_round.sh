#!/bin/bash
# One GPU call: parity tests, bench (+ reference arm), phase breakdown, ncu launch list + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
if [ -f uncalled_b200/libunc_b200_pt.so ]; then timeout 300 python tools/gpu_phases.py g4m7 2368 > gpurun_out/phases.txt 2>&1; tail -14 gpurun_out/phases.txt; fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --reads 2368 > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_map -s 1 -c 1 -o gpurun_out/k2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --reads 2368 > gpurun_out/k2_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_events -s 1 -c 1 -o gpurun_out/k1_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --reads 10000 > gpurun_out/k1_full.log 2>&1
ls -la gpurun_out
