#!/bin/bash
# level-1 whole-window rounds, second version (static / dynamic steps of the window walk): parity with the switch on, the
# -DBDF_CHECK build, throughput with the bucket prefetch (window bits 2 / 4) and with 10 / 12 resident CTAs per SM
# (-DBDF_L1_MIN_CTAS builds: 48 / 40 registers)
mkdir -p gpurun_out
L=$PWD/libdeflate_rsx_b200
(VARIANTS=0:8,1:8,3:8,5:8 timeout 300 python -u gpurun_scripts/l1_window_probe.py 8192
 echo "== libbdeflate_c10.so"; BDF_LIBRARY=$L/libbdeflate_c10.so VARIANTS=1:10,3:10,1:8 timeout 300 python -u gpurun_scripts/l1_window_probe.py 8192 text,mixedB,binary,lowent
 echo "== libbdeflate_c12.so"; BDF_LIBRARY=$L/libbdeflate_c12.so VARIANTS=1:12,3:12,1:10 timeout 300 python -u gpurun_scripts/l1_window_probe.py 8192 text,mixedB,binary,lowent) 2>&1 | tee gpurun_out/l1_window_probe_r4b.txt | grep -v "^$" | tail -16
BDF_L1_WINDOW=1 timeout 500 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py tests/test_gpu_device_any.py tests/test_gpu_host_paths.py -x -q -k "not near_optimal and not ratio_tier and not decompress and not config4 and not config5 and not checksum" 2>&1 | tail -4
BDF_L1_WINDOW=3 BDF_LIBRARY=$L/libbdeflate_check.so timeout 300 python tests/check_build_helper.py 2>&1 | tail -2
