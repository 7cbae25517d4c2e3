#!/bin/bash
# First GPU call of the next round (run `bash tools/build_variants.sh` on the CPU box first):
#   1. parity tests            2. every configuration in uncalled_b200/variants/ on the bench workload
#   3. per-phase cycles (two observers) of the shipped build and of the lean 16-warp build
#   4. the chunk-streaming workload (configs[4]-like), which has not run on hardware yet
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/gpu_variants.py 10000 2>&1 | tee gpurun_out/variants_next.txt
timeout 200 python tools/gpu_phases.py g4m7 2368 > gpurun_out/phases_default.txt 2>&1; tail -24 gpurun_out/phases_default.txt
for v in uncalled_b200/variants_pt/*.so; do
  [ -f "$v" ] || continue
  timeout 200 python tools/gpu_phases.py g4m7 2368 "$v" > "gpurun_out/phases_$(basename "$v" .so).txt" 2>&1; tail -24 "gpurun_out/phases_$(basename "$v" .so).txt"
done
timeout 600 python bench.py --workload stream --steps 2 --warmup 1 > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err; echo "stream bench rc=$?"; cut -c1-600 gpurun_out/bench_stream.json; tail -3 gpurun_out/bench_stream.err
# batch tail: two pools used alternately (expected from a list-scheduling simulation of the oracle's per-read costs:
# makespan / ideal = 1.07 at 2 CTAs/SM, 1.11 at 3, 1.15 at 4 for 10 000-read batches)
timeout 600 python bench.py --overlap --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/bench_overlap.json 2> gpurun_out/bench_overlap.err; echo "overlap bench rc=$?"; cut -c1-900 gpurun_out/bench_overlap.json
# ordered mode (`uncalled map -t 1` semantics, unc_map_batch_ordered): cost of the re-mapping rounds on the bench workload
timeout 600 python bench.py --ordered --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_ordered.json 2> gpurun_out/bench_ordered.err; echo "ordered bench rc=$?"; python -c "import json;print(json.load(open('gpurun_out/bench_ordered.json'))['ordered'])"
# exact-ties kernel (unc_pool_set_tie_order(1): the reference's pdqsort reproduced): its cost on the bench workload
timeout 900 python bench.py --exact-ties --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err; echo "exact-ties bench rc=$?"; python -c "import json;print(json.load(open('gpurun_out/bench_exact.json'))['exact_ties'])"
