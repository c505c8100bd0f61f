#!/bin/bash
# call 21: litlen table in shared memory + offset table in global memory (BDF_LANE_CFG=6 / 7)
mkdir -p gpurun_out
PRODUCERS=1 KINDS=text,binary,mixedB,lowent timeout 600 python -u gpurun_scripts/inflate_modes.py 65536 lane0 lane5 lane6 lane7 2>&1 | tee gpurun_out/inflate_modes_r3u.txt | tail -5
