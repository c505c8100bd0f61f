#!/bin/bash
# lane-per-stream inflate with cache hints: experiment builds libbdeflate_h<mask>.so with the table look-ups as
# ld.global.L1::evict_last (1), the compressed input as ld.global.nc.L1::no_allocate (2), the output sectors as st.global.cs (4).
# All three measured slower than plain loads / stores (profiles/r4_lane_hints.txt), so the macro was taken out of inflate_lane.cuh again.
export PRODUCERS=1 KINDS=${KINDS:-text,binary,mixedB}
for h in "" _h7 _h1 _h6; do
  echo "== libbdeflate$h.so"
  BDF_LIBRARY=$PWD/libdeflate_rsx_b200/libbdeflate$h.so timeout 200 python -u gpurun_scripts/inflate_modes.py 65536 lane5 auto 2>&1 | tail -4
done
