#!/bin/bash
# diagnostics: EM throughput vs number of panels (tail / contention) and vs CTA stagger
cd /root/repo
for P in 148 296 592 1184 1250 1480; do
  timeout 120 python bench.py --no-cpu --steps 3 --warmup 3 --panels $P 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('panels', d['config'].get('panels_per_gpu'), 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'kernel_ms', d['roofline']['kernel_ms'])"
done
for S in 40000 80000 145000 200000; do
  DFM_FUSED_STAGGER=$S timeout 120 python bench.py --no-cpu --steps 3 --warmup 3 --panels 1184 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stagger $S panels 1184 value', round(d['value']), 'ms', round(d['ms_per_step'],3))"
done
