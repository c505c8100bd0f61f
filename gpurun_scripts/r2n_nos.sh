#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -u gpurun_scripts/nos_probe.py 2>&1 | tee gpurun_out/nos_probe_$1.txt | tail -16
timeout 1500 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_fuzz.py tests/test_gpu_host_paths.py tests/test_gpu_guard.py tests/test_gpu_size.py tests/test_gpu_configs.py -x -q 2>&1 | tail -12
