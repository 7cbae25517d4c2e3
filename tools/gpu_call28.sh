#!/bin/bash
# final check of the tree as committed: GPU test suite, smoke(), a short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu28.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu28.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/bench28.json 2> gpurun_out/bench28.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench28.json
