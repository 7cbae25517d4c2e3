#!/bin/bash
# Round 2, final 1-GPU validation of the kernel that ships (second structure + faster seed tracker): parity suite, headline bench,
# reference arm, streaming and fast5 workloads, phases, launch list, DRAM traffic of the full launch, ncu --set full.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu23.log
tail -4 gpurun_out/pytest_gpu23.log
timeout 900 python bench.py > gpurun_out/bench23.json 2> gpurun_out/bench23.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench23.json; tail -3 gpurun_out/bench23.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench23_ref.json 2> gpurun_out/bench23_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench23_ref.json
timeout 600 python bench.py --workload stream --steps 2 --warmup 1 > gpurun_out/bench23_stream.json 2> gpurun_out/bench23_stream.err; echo "stream rc=$?"; cut -c1-1500 gpurun_out/bench23_stream.json; tail -3 gpurun_out/bench23_stream.err
timeout 600 python bench.py --workload fast5 --files 16 --steps 1 --warmup 1 > gpurun_out/bench23_fast5.json 2> gpurun_out/bench23_fast5.err; echo "fast5 rc=$?"; cut -c1-800 gpurun_out/bench23_fast5.json; tail -3 gpurun_out/bench23_fast5.err
for v in uncalled_b200/variants_pt/*.so; do
  timeout 200 python tools/gpu_phases.py g4m7 2368 "$v" > "gpurun_out/phases23_$(basename "$v" .so).txt" 2>&1; tail -12 "gpurun_out/phases23_$(basename "$v" .so).txt"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches23.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/b_ncu23.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k2_map -s 1 -c 1 --csv --log-file gpurun_out/k2_traffic23.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --reads 10000 > gpurun_out/k2_traffic23.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_map -s 1 -c 1 -o gpurun_out/k2_r2b_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --reads 1184 > gpurun_out/k2_r2b_full.log 2>&1
ls -la gpurun_out | tail -12
