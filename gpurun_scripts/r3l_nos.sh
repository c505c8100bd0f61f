#!/bin/bash
# call 13: near-optimal tier as three kernels per wave (deflate_nos_split.cuh): parity first, then throughput A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_configs.py tests/test_gpu_determinism.py tests/test_gpu_check_build.py -x -q 2>&1 | tail -5
timeout 600 python -u gpurun_scripts/nos_probe.py 2048 2>&1 | tee gpurun_out/nos_probe_r3l_split.txt | tail -16
BDF_NOS_SPLIT=0 timeout 600 python -u gpurun_scripts/nos_probe.py 2048 2>&1 | grep "L12" | tee gpurun_out/nos_probe_r3l_single.txt | tail -6
