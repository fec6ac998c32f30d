#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q -k "em_ or c3 or nile or frozen" 2>&1 | tail -3 > gpurun_out/q8.log
for c in c3; do
  timeout 300 python bench.py --config $c --steps 3 --warmup 2 2>gpurun_out/bench_$c.err | tail -1 > gpurun_out/bench_$c.json
  python -c "
import json; d=json.loads(open('gpurun_out/bench_$c.json').read().strip().splitlines()[-1]); print('$c', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['roofline']['kernel_ms'])" >> gpurun_out/q8.log 2>&1
  tail -2 gpurun_out/bench_$c.err >> gpurun_out/q8.log
done
timeout 300 python tools/bench_c1_em.py 296 5 4 >> gpurun_out/q8.log 2>&1
timeout 300 python tools/bench_c1_em.py 1000 10 4 >> gpurun_out/q8.log 2>&1
cat gpurun_out/q8.log
