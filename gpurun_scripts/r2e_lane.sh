#!/bin/bash
# lane engine at full occupancy (65536 streams) and with capped warps per SM
mkdir -p gpurun_out
export PRODUCERS=1
timeout 600 python -u gpurun_scripts/inflate_modes.py 65536 group lane0 lane1 auto 2>&1 | tee gpurun_out/inflate_modes_$1.txt | tail -8
for w in 3 5; do
  echo "== BDF_LANE_WARPS=$w"
  BDF_LANE_WARPS=$w KINDS=text,binary timeout 300 python -u gpurun_scripts/inflate_modes.py 65536 lane0 2>&1 | tail -2
done
