#!/bin/bash
# quick GPU check of the fused kernels: parity tests, bench at a balanced and at the C5 shard size, phase timers
cd /root/repo
timeout 400 python -m pytest tests -m gpu -x -q -k "fused or als or em" 2>&1 | tail -3
for P in 1184 1250; do
  timeout 120 python bench.py --no-cpu --steps 3 --warmup 3 --panels $P 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('panels $P value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), 'als', round(d['als']['value']))"
done
DFM_FUSED_PHASES=1 timeout 150 python bench.py --no-cpu --steps 2 --warmup 3 --panels 1184 2>&1 >/dev/null | head -17
