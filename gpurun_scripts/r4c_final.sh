#!/bin/bash
# end of round 2: level-1 stress (whole-window rounds, also under -DBDF_CHECK), whole GPU suite, smoke, both bench arms,
# launch list of the bench command, ncu capture of the level-1 kernel on text as shipped
TAG=${1:-r4}
mkdir -p gpurun_out
L=$PWD/libdeflate_rsx_b200
(timeout 300 python -u gpurun_scripts/stress_deflate_l1.py 30000 1
 BDF_LIBRARY=$L/libbdeflate_check.so STRESS_FORMATS=2 timeout 300 python -u gpurun_scripts/stress_deflate_l1.py 12000 2) 2>&1 | tee gpurun_out/stress_l1_$TAG.txt | tail -9
timeout 1000 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
timeout 400 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -2 gpurun_out/bench_${TAG}.err
cut -c1-200 gpurun_out/bench_${TAG}.json
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"], 1))
for k in ("compress_l1", "compress_l6", "compress_l12", "mixed_pipeline"):
    if k in d: print(k, round(d[k]["value"], 2), {kk: round(vv["value"], 2) for kk, vv in d[k].items() if isinstance(vv, dict) and "value" in vv})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --pipeline-streams 4096 > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_launches_${TAG}.log | cut -c1-200
export PROFILE_OUT=gpurun_out/profiles_r4
mkdir -p $PROFILE_OUT
NCU="ncu --set full --clock-control none --import-source on -s 1 -c 1 -f"
name=r4_l1_text
python -u gpurun_scripts/deflate_probe.py 1 8192 text > gpurun_out/plain_$name.log 2>&1 && \
  timeout 400 $NCU -k regex:deflate_l1 -o gpurun_out/prof_$name python -u gpurun_scripts/deflate_probe.py 1 8192 text > gpurun_out/ncu_$name.log 2>&1
tail -1 gpurun_out/ncu_$name.log | cut -c1-120
python tools/profile_summary.py gpurun_out/prof_$name.ncu-rep deflate_l1_kernel $name 8192 88500 deflate_l1_text_window libdeflate_rsx_b200/csrc/deflate_l1.cuh "$(grep -h 'GB/s' gpurun_out/plain_$name.log | tail -1 | cut -c1-150)" > /dev/null 2> gpurun_out/sum_$name.err
rm -f gpurun_out/prof_$name.ncu-rep
ls $PROFILE_OUT; tail -3 gpurun_out/sum_$name.err
