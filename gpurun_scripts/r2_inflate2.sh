#!/bin/bash
# inflate iteration with the wider parity net: inflate + fuzz + config tests, bench at G=16/32, text throughput
TAG=${1:-r2i}
timeout 400 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_fuzz.py tests/test_gpu_api_stream.py -x -q 2>&1 | tail -4
for g in 16 32; do
  export BDF_INFLATE_GROUP=$g
  echo "== G=$g"
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipeline-streams 8192 2>gpurun_out/bench_${TAG}_g$g.err | tee gpurun_out/bench_${TAG}_g$g.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('corpusA value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), 'mixed decompress', round(d['mixed_pipeline']['decompress'],1))"
  timeout 90 python -u gpurun_scripts/gpu_quick.py 2>&1 | tail -1
done
unset BDF_INFLATE_GROUP
if [ "$2" != "noncu" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:inflate_kernel -s 3 -c 1 -f -o gpurun_out/prof_inflate_${TAG} python bench.py --streams 16384 --steps 3 --warmup 3 --no-cpu-baseline --pipeline-streams 0 > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log | cut -c1-200
fi
