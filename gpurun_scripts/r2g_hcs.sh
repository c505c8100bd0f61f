#!/bin/bash
# new shared-memory hash-chain kernel: parity tests, then throughput against the old kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_fuzz.py tests/test_gpu_size.py tests/test_gpu_host_paths.py tests/test_gpu_configs.py -x -q 2>&1 | tail -15
timeout 600 python -u gpurun_scripts/deflate_probe.py 2,6,9 8192 2>&1 | tee gpurun_out/deflate_probe_$1.txt | tail -20
echo "== old kernel"
BDF_HC_KERNEL=old timeout 600 python -u gpurun_scripts/deflate_probe.py 6 8192 2>&1 | tail -6
