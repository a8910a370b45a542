#!/bin/bash
for b in 32 64 96 128 192 100000; do
EMF_RAY_BUDGET=$b timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('budget $b', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stages_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],4))"
done
