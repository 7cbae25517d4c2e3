#!/bin/bash
# Round 2, seventh GPU call: the kernel that ships (second structure, count / prefix / emit phases) -- parity suite,
# CTA shapes, phases, headline bench, streaming and fast5 workloads, launch list, ncu --set full, DRAM traffic of the full launch.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu7.log
tail -4 gpurun_out/pytest_gpu7.log
python tools/gpu_variants.py 10000 2>&1 | tee gpurun_out/variants7.txt
for v in uncalled_b200/variants_pt/*.so; do
  timeout 200 python tools/gpu_phases.py g4m7 2368 "$v" > "gpurun_out/phases7_$(basename "$v" .so).txt" 2>&1; tail -15 "gpurun_out/phases7_$(basename "$v" .so).txt"
done
timeout 900 python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/bench7.json; tail -3 gpurun_out/bench7.err
timeout 600 python bench.py --workload stream --steps 2 --warmup 1 > gpurun_out/bench7_stream.json 2> gpurun_out/bench7_stream.err; echo "stream rc=$?"; cut -c1-1500 gpurun_out/bench7_stream.json; tail -3 gpurun_out/bench7_stream.err
timeout 600 python bench.py --workload fast5 --files 16 --steps 1 --warmup 1 > gpurun_out/bench7_fast5.json 2> gpurun_out/bench7_fast5.err; echo "fast5 rc=$?"; cut -c1-800 gpurun_out/bench7_fast5.json; tail -3 gpurun_out/bench7_fast5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k2_map -s 1 -c 1 --csv --log-file gpurun_out/k2_traffic_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --reads 10000 > gpurun_out/k2_traffic_r2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_map -s 1 -c 1 -o gpurun_out/k2_r2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --reads 1184 > gpurun_out/k2_r2_full.log 2>&1
ls -la gpurun_out | tail -8
