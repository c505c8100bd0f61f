#!/bin/bash
# lane-per-stream inflate engine: parity in every mode, then throughput per corpus kind
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engines.py -x -q 2>&1 | tail -15
timeout 400 python -u gpurun_scripts/inflate_modes.py 16384 2>&1 | tee gpurun_out/inflate_modes_${1:-r2b}.txt | tail -12
