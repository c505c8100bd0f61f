#!/bin/bash
# ncu --set full of the hash-chain compressor (level 6): launch 3 = corpus A, launch 7 = mixed corpus
timeout 200 ncu --set full --clock-control none --import-source on -k regex:deflate_hc_kernel -s 2 -c 1 -f -o gpurun_out/prof_hc_a_r3 python gpurun_scripts/gpu_compress.py 6 4736 > gpurun_out/prof_hc_a_r3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:deflate_hc_kernel -s 6 -c 1 -f -o gpurun_out/prof_hc_m_r3 python gpurun_scripts/gpu_compress.py 6 4736 > gpurun_out/prof_hc_m_r3.log 2>&1
tail -2 gpurun_out/prof_hc_a_r3.log gpurun_out/prof_hc_m_r3.log
