#!/bin/bash
# DTW on the GPU: parity tests through the C-ABI / the Python classes / the _uncalled module, then a timing next to the reference's DTW.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dtw.py tests/test_pymodule.py -m gpu -x -q > gpurun_out/pytest_dtw27.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_dtw27.log
timeout 300 python tools/bench_dtw.py > gpurun_out/bench_dtw27.json 2> gpurun_out/bench_dtw27.err; echo "bench rc=$?"; cat gpurun_out/bench_dtw27.json; tail -3 gpurun_out/bench_dtw27.err
timeout 300 python tools/bench_dtw.py --problems 2048 --kmers 300 --events 450 --cpu-problems 512 >> gpurun_out/bench_dtw27.json 2>> gpurun_out/bench_dtw27.err; tail -1 gpurun_out/bench_dtw27.json
