#!/bin/bash
cd /root/repo
bash tools/sweep_variants.sh > gpurun_out/q6.log 2>&1
for c in c4 c2-single c3; do
  timeout 300 python bench.py --config $c --steps 3 --warmup 2 2>gpurun_out/bench_$c.err | tail -1 > gpurun_out/bench_$c.json
  tail -c 600 gpurun_out/bench_$c.json >> gpurun_out/q6.log; echo >> gpurun_out/q6.log; tail -3 gpurun_out/bench_$c.err >> gpurun_out/q6.log
done
timeout 300 python tools/bench_c1_em.py 296 5 4 >> gpurun_out/q6.log 2>&1
cat gpurun_out/q6.log
