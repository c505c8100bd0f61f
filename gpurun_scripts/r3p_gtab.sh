#!/bin/bash
# call 17: lane-per-stream inflate with the per-lane tables in global memory (read through L1): 16 warps per SM
mkdir -p gpurun_out
PRODUCERS=1 KINDS=text,binary,mixedB,lowent timeout 600 python -u gpurun_scripts/inflate_modes.py 65536 lane0 lane3 lane4 lane3w12 lane3w10 2>&1 | tee gpurun_out/inflate_modes_r3p.txt | tail -5
