#!/bin/bash
# round 2, last GPU calls: level-1 whole-window rounds (throughput against the one-match rounds, resident CTAs; parity with the
# switch on; the -DBDF_CHECK build), the lane-kernel cache hints, and an ncu capture of the new level-1 rounds on text
mkdir -p gpurun_out
timeout 300 python -u gpurun_scripts/l1_window_probe.py 8192 2>&1 | tee gpurun_out/l1_window_probe_r4a.txt | tail -8
BDF_L1_WINDOW=1 timeout 500 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py tests/test_gpu_device_any.py tests/test_gpu_host_paths.py -x -q -k "not near_optimal and not ratio_tier and not decompress and not config4 and not config5 and not checksum" 2>&1 | tail -4
BDF_L1_WINDOW=1 BDF_LIBRARY=$PWD/libdeflate_rsx_b200/libbdeflate_check.so timeout 300 python tests/check_build_helper.py 2>&1 | tail -2
bash gpurun_scripts/lane_hints_probe.sh 2>&1 | tee gpurun_out/lane_hints_r4a.txt | grep -v "^$" | tail -16
export PROFILE_OUT=gpurun_out/profiles_r4
mkdir -p $PROFILE_OUT
NCU="ncu --set full --clock-control none --import-source on -s 1 -c 1 -f"
name=r4_l1_text_window
BDF_L1_WINDOW=1 python -u gpurun_scripts/deflate_probe.py 1 8192 text > gpurun_out/plain_$name.log 2>&1 && \
  BDF_L1_WINDOW=1 timeout 400 $NCU -k regex:deflate_l1 -o gpurun_out/prof_$name python -u gpurun_scripts/deflate_probe.py 1 8192 text > gpurun_out/ncu_$name.log 2>&1
tail -1 gpurun_out/ncu_$name.log | cut -c1-120
python tools/profile_summary.py gpurun_out/prof_$name.ncu-rep deflate_l1_kernel $name 8192 88500 deflate_l1_text_window libdeflate_rsx_b200/csrc/deflate_l1.cuh "$(grep -h 'GB/s' gpurun_out/plain_$name.log | tail -1 | cut -c1-150)" > /dev/null 2> gpurun_out/sum_$name.err
rm -f gpurun_out/prof_$name.ncu-rep
ls $PROFILE_OUT; tail -3 gpurun_out/sum_$name.err
