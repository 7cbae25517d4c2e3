#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_variants.py 10000 2>&1 | tee gpurun_out/variants6.txt
for v in uncalled_b200/variants_pt/*.so; do
  timeout 200 python tools/gpu_phases.py g4m7 2368 "$v" > "gpurun_out/phases6_$(basename "$v" .so).txt" 2>&1; tail -14 "gpurun_out/phases6_$(basename "$v" .so).txt"
done
timeout 600 python -m pytest tests/test_pymodule.py tests/test_z_gpu_api_and_tools.py -m gpu -x -q -k "pymodule or map_pool or timeout or stream" > gpurun_out/pytest_gpu6.log 2>&1; tail -4 gpurun_out/pytest_gpu6.log
timeout 900 python bench.py --workload stream --steps 2 --warmup 1 > gpurun_out/bench6_stream.json 2> gpurun_out/bench6_stream.err; echo "stream rc=$?"; cut -c1-2500 gpurun_out/bench6_stream.json; tail -3 gpurun_out/bench6_stream.err
timeout 900 python bench.py --workload fast5 --files 16 --steps 1 --warmup 1 > gpurun_out/bench6_fast5.json 2> gpurun_out/bench6_fast5.err; echo "fast5 rc=$?"; cut -c1-2000 gpurun_out/bench6_fast5.json; tail -3 gpurun_out/bench6_fast5.err
