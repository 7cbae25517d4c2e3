#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_variants.py 10000 2>&1 | tee gpurun_out/variants4.txt
