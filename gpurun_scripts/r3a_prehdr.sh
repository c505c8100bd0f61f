#!/bin/bash
# round 2, session 2, call 1: header pre-pass (inflate_prehdr.cuh) + 4-byte tail reject in the chain walk
# whole GPU suite first (parity), then A/B of both changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
PRODUCERS=1 KINDS=corpusA,text,binary,mixedB timeout 500 python -u gpurun_scripts/inflate_modes.py 65536 auto auto_nopre lane0 lane0_nopre 2>&1 | tee gpurun_out/inflate_modes_r3a.txt | tail -6
timeout 400 python -u gpurun_scripts/deflate_probe.py 2,6,9 8192 text,binary,mixedB 2>&1 | tee gpurun_out/deflate_probe_r3a_tail4.txt | tail -10
BDF_LIBRARY=$PWD/libdeflate_rsx_b200/libbdeflate_tail0.so timeout 400 python -u gpurun_scripts/deflate_probe.py 2,6,9 8192 text,binary,mixedB 2>&1 | tee gpurun_out/deflate_probe_r3a_tail0.txt | tail -10
