#!/bin/bash
# Round 2, eighth GPU call: configs[3] -- chr1-sized index built on the box, long reads, max_events 30000.
mkdir -p gpurun_out
timeout 1700 python tools/bench_chr1.py --mb 230 --reads 2048 --samples 32000 --keep > gpurun_out/bench_chr1.json 2> gpurun_out/bench_chr1.err; echo "chr1 rc=$?"; cut -c1-3000 gpurun_out/bench_chr1.json; tail -5 gpurun_out/bench_chr1.err
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k2_map -s 1 -c 1 --csv --log-file gpurun_out/k2_chr1_metrics.csv python tools/bench_chr1.py --mb 230 --reads 2048 --samples 32000 --reuse --no-cpu --steps 1 > gpurun_out/k2_chr1_ncu.log 2>&1
tail -8 gpurun_out/k2_chr1_metrics.csv | cut -c1-60,330-
