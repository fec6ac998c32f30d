#!/bin/bash
# round-2 evidence: gpu test tier, bench lines of every config, ncu launch list, one full ncu capture of the EM kernel
cd /root/repo
tag=${1:-r02a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log
timeout 300 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 300 gpurun_out/bench_$tag.json; echo
for c in c4 c3 c2-single; do
  timeout 300 python bench.py --config $c > gpurun_out/bench_${c}_$tag.json 2> gpurun_out/bench_${c}_$tag.err; tail -c 300 gpurun_out/bench_${c}_$tag.json; echo
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; tail -c 300 gpurun_out/bench_ref_$tag.json; echo
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_l_$tag.log 2>&1; tail -1 gpurun_out/ncu_l_$tag.log | cut -c1-100
