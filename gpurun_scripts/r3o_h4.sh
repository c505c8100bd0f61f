#!/bin/bash
# call 16: near-optimal tier on 4-byte hash chains + one 3-byte candidate (libbdeflate_h4.so): size and speed by depth
mkdir -p gpurun_out
export BDF_LIBRARY=$PWD/libdeflate_rsx_b200/libbdeflate_h4.so
for d in 30 60 120 250 500; do
  BDF_NOS_DEPTH=$d KINDS=text,mixedB,binary,lowent LEVELS=12 timeout 300 python -u gpurun_scripts/nos_probe.py 2048 2>&1 | sed "s/^/h4 depth $d /" | tee -a gpurun_out/nos_probe_r3o_h4.txt | tail -4
done
unset BDF_LIBRARY
for d in 150 300; do
  BDF_NOS_DEPTH=$d KINDS=text,mixedB LEVELS=12 timeout 300 python -u gpurun_scripts/nos_probe.py 2048 2>&1 | sed "s/^/h3 depth $d /" | tee -a gpurun_out/nos_probe_r3o_h4.txt | tail -2
done
