#!/bin/bash
# Round 2, last call: configs[3]-like chr1-sized index with the kernel that ships; ncu --set full of the other mapper kernels
# (exact ties, ordered, streaming) and of the self-alignment kernels of `uncalled index`.
mkdir -p gpurun_out
timeout 900 python tools/bench_chr1.py --mb 230 --reads 2048 --samples 32000 --steps 2 > gpurun_out/bench25_chr1.json 2> gpurun_out/bench25_chr1.err; echo "chr1 rc=$?"; cut -c1-1800 gpurun_out/bench25_chr1.json; tail -3 gpurun_out/bench25_chr1.err
timeout 300 ncu --set full --clock-control none -k regex:"k2_map_(exact|ord)" -c 2 -o gpurun_out/k2_exact_ord_r2 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --reads 592 > gpurun_out/k2_exact_ord_r2.log 2>&1; echo "ncu exact/ord rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:k2_map_stream -s 4 -c 1 -o gpurun_out/k2_stream_r2 -f python bench.py --workload stream --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/k2_stream_r2.log 2>&1; echo "ncu stream rc=$?"
rm -f /tmp/sa.*; cp bench_data/g4m7.fa /tmp/sa.fa
timeout 300 ncu --set full --clock-control none -k regex:k_selfalign -c 2 -o gpurun_out/k_selfalign_r2 -f python -m uncalled_b200 index /tmp/sa.fa > gpurun_out/k_selfalign_r2.log 2>&1; echo "ncu selfalign rc=$?"
ls -la gpurun_out | tail -8
