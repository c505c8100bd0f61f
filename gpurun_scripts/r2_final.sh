#!/bin/bash
# end-of-iteration evidence: whole GPU test suite, smoke, bench line (both arms), ncu launch list of the bench command
TAG=${1:-r2}
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
timeout 400 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -2 gpurun_out/bench_${TAG}.err
cut -c1-200 gpurun_out/bench_${TAG}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --pipeline-streams 4096 > gpurun_out/ncu_launches_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launches_${TAG}.log | cut -c1-200
