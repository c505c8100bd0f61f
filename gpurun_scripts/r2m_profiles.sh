#!/bin/bash
# ncu --set full captures of the dominant kernels (each after its plain run); summarised on the box
# (the reports together exceed what gpurun brings back), two reports kept
mkdir -p gpurun_out/profiles_r2
export PROFILE_OUT=gpurun_out/profiles_r2
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
KINDS=mixedB cap r2_lane_mixed inflate_lane inflate_lane_kernel 49152 81700 inflate_corpusB $C/inflate_lane.cuh 1 python -u gpurun_scripts/inflate_modes.py 65536 auto
KINDS=text cap r2_lane_text inflate_lane inflate_lane_kernel 65536 80970 inflate_text $C/inflate_lane.cuh 0 python -u gpurun_scripts/inflate_modes.py 65536 lane0
KINDS=corpusA cap r2_group_corpusA 'inflate_kernel' 'inflate_kernel' 16384 65936 inflate_config2 $C/inflate.cuh 0 python -u gpurun_scripts/inflate_modes.py 16384 group
cap r2_hcs_mixed deflate_hcs deflate_hcs_kernel 1536 81700 deflate_l6_corpusB $C/deflate_hcs.cuh 1 python -u gpurun_scripts/deflate_probe.py 6 2048 mixedB
cap r2_hcs_text_l2 deflate_hcs deflate_hcs_kernel 1216 82430 deflate_l2_text $C/deflate_hcs.cuh 0 python -u gpurun_scripts/deflate_probe.py 2 1216 text
cap r2_hc_corpusA 'deflate_hc_kernel' deflate_hc_kernel 8192 65930 deflate_l6_corpusA $C/deflate_hc.cuh 0 python -u gpurun_scripts/deflate_probe.py 6 8192 corpusA
cap r2_l1_corpusA deflate_l1 deflate_l1_kernel 16384 66180 deflate_l1_corpusA $C/deflate_l1.cuh 0 python -u gpurun_scripts/deflate_probe.py 1 16384 corpusA
cap r2_l1_text deflate_l1 deflate_l1_kernel 8192 89000 deflate_l1_text $C/deflate_l1.cuh 0 python -u gpurun_scripts/deflate_probe.py 1 8192 text
ls -la gpurun_out/profiles_r2 | tail -12; cat gpurun_out/sum_*.err | tail -5
