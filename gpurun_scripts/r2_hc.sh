#!/bin/bash
# hash-chain compressor iteration: parity first on small batches with tight timeouts (a pipeline bug hangs), then throughput
timeout 120 python -m pytest tests/test_gpu_checksum_compress.py -x -q -k "byte_identical_to_oracle or failure_is_in_band" 2>&1 | tail -3 || exit 1
timeout 300 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py -x -q 2>&1 | tail -3
timeout 300 python gpurun_scripts/gpu_compress.py 2,6 8192 2>&1 | grep "L[0-9]:" 
