#!/bin/bash
# hash-chain compressor iteration: parity (compress + fuzz tests), then throughput at levels 2/6/9
timeout 600 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -3
timeout 400 python gpurun_scripts/gpu_compress.py 2,6 8192 2>&1 | grep "L[0-9]:" 
