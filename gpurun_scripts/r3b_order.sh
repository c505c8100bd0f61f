#!/bin/bash
# call 2: is "auto + pre-pass" really slower on binary / mixed data, or was it the first position of the
# measurement?  (same modes, other order, twice); then ncu of the lane-group kernel on config 2 with the pre-pass
mkdir -p gpurun_out/profiles_r3
PRODUCERS=1 KINDS=binary,mixedB timeout 300 python -u gpurun_scripts/inflate_modes.py 65536 auto_nopre auto lane0 auto auto_nopre 2>&1 | tee gpurun_out/inflate_modes_r3b.txt | tail -4
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
KINDS=corpusA cap r3_group_corpusA 'inflate_kernel' 'inflate_kernel' 16384 65936 inflate_config2 $C/inflate.cuh 0 python -u gpurun_scripts/inflate_modes.py 16384 group
ls -la gpurun_out/profiles_r3 | tail; cat gpurun_out/sum_*.err | tail -5
