#!/bin/bash
timeout 150 python -m pytest tests/test_gpu_checksum_compress.py -x -q -k "byte_identical_to_oracle or failure_is_in_band" 2>&1 | tail -2 || exit 1
timeout 400 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py tests/test_gpu_guard.py -x -q 2>&1 | tail -3
timeout 300 python gpurun_scripts/gpu_compress.py 1 16384 2>&1 | grep "L1:"
