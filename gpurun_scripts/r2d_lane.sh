#!/bin/bash
# lane engine: parity in every mode, throughput per corpus kind, then one ncu capture on text
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engines.py -x -q 2>&1 | tail -15
timeout 400 python -u gpurun_scripts/inflate_modes.py 16384 2>&1 | tee gpurun_out/inflate_modes_$1.txt | tail -12
export KINDS=text PRODUCERS=1
timeout 120 python -u gpurun_scripts/inflate_modes.py 8192 lane0 > gpurun_out/plain_$1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inflate_lane -s 1 -c 1 -f -o gpurun_out/prof_lane_$1 python -u gpurun_scripts/inflate_modes.py 8192 lane0 > gpurun_out/ncu_$1.log 2>&1
tail -3 gpurun_out/ncu_$1.log | cut -c1-200
