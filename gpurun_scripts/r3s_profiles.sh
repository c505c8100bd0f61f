#!/bin/bash
# ncu --set full captures of the kernels that changed in the second half of round 2 (summarised on the box)
mkdir -p gpurun_out/profiles_r3
export PROFILE_OUT=gpurun_out/profiles_r3
NCU="ncu --set full --clock-control none --import-source on -s 1 -c 1 -f"
cap() {   # name, kernel regex, kernel substr, nstreams, algorithmic bytes per stream, traffic key, source, keep, command...
  local name=$1 rx=$2 sub=$3 ns=$4 alg=$5 key=$6 src=$7 keep=$8; shift 8
  "$@" > gpurun_out/plain_$name.log 2>&1 && timeout 900 $NCU -k regex:$rx -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name: $(tail -1 gpurun_out/ncu_$name.log | cut -c1-100)"
  python tools/profile_summary.py gpurun_out/prof_$name.ncu-rep "$sub" $name $ns $alg $key $src "$(grep -h 'GB/s' gpurun_out/plain_$name.log | tail -1 | cut -c1-150)" > /dev/null 2> gpurun_out/sum_$name.err
  [ "$keep" = 1 ] || rm -f gpurun_out/prof_$name.ncu-rep
}
export PRODUCERS=1
C=libdeflate_rsx_b200/csrc
KINDS=mixedB cap r3_lane_mixed inflate_lane inflate_lane_kernel 49152 81700 inflate_corpusB $C/inflate_lane.cuh 0 python -u gpurun_scripts/inflate_modes.py 65536 auto
KINDS=text cap r3_lane_text inflate_lane inflate_lane_kernel 65536 80970 inflate_text $C/inflate_lane.cuh 0 python -u gpurun_scripts/inflate_modes.py 65536 lane5
cap r3_hcs_mixed deflate_hcs deflate_hcs_kernel 1536 81700 deflate_l6_corpusB $C/deflate_hcs.cuh 0 python -u gpurun_scripts/deflate_probe.py 6 2048 mixedB
KINDS=corpusA LEVELS=12 cap r3_nos_cost_corpusA deflate_nos_cost deflate_nos_cost_kernel 2048 65930 deflate_l12_corpusA $C/deflate_nos_split.cuh 0 python -u gpurun_scripts/nos_probe.py 2048
ls -la gpurun_out/profiles_r3 | tail -8; cat gpurun_out/sum_*.err | tail -5
