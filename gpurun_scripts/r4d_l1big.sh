#!/bin/bash
# level 1, units above 64 KiB and the size estimator with whole-window rounds (BDF_L1_WINDOW bit 3): throughput and parity
mkdir -p gpurun_out
L=$PWD/libdeflate_rsx_b200
timeout 300 python -u gpurun_scripts/l1_big_probe.py 1024 2>&1 | tee gpurun_out/l1_big_probe_r4d.txt | tail -5
BDF_L1_WINDOW=11 timeout 600 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_fuzz.py tests/test_gpu_size.py tests/test_gpu_device_any.py tests/test_gpu_api_stream.py tests/test_gpu_host_paths.py -x -q -k "not near_optimal and not ratio_tier and not decompress and not checksum_reference and not checksum_tails" 2>&1 | tail -4
BDF_L1_WINDOW=11 BDF_LIBRARY=$L/libbdeflate_check.so STRESS_FORMATS=0 timeout 300 python -u gpurun_scripts/stress_deflate_l1.py 12000 3 2>&1 | tail -3
