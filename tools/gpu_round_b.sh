#!/bin/bash
# Final call of a round: parity tests, bench, DRAM traffic of the full-size k2_map launch, launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 300 python tools/gpu_phases.py g4m7 2368 > gpurun_out/phases.txt 2>&1; tail -12 gpurun_out/phases.txt
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k2_map -s 1 -c 1 --csv --log-file gpurun_out/k2_traffic_full.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --reads 10000 > gpurun_out/k2_traffic_full.log 2>&1
tail -4 gpurun_out/k2_traffic_full.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
