#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/q7.log
bash tools/sweep_variants.sh >> gpurun_out/q7.log 2>&1
timeout 200 python bench.py --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MAIN', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), 'ok', d['config']['all_status_ok'])" >> gpurun_out/q7.log
for c in c4 c3; do
  timeout 300 python bench.py --config $c --steps 3 --warmup 2 2>gpurun_out/bench_$c.err | tail -1 > gpurun_out/bench_$c.json
  python -c "
import json; d=json.loads(open('gpurun_out/bench_$c.json').read().strip().splitlines()[-1]); print('$c', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['roofline']['kernel_ms'])" >> gpurun_out/q7.log 2>&1
  tail -2 gpurun_out/bench_$c.err >> gpurun_out/q7.log
done
cat gpurun_out/q7.log
