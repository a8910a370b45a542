#!/bin/bash
# bench.py once per environment setting: scripts/env_ab.sh "EMF_RAY_SCHED=0" "EMF_RAY_FRONT=10" ...
for t in "$@"; do
  env $t timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$t', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stages_ms'].items() if isinstance(v,(int,float))}, 'e2e', d['e2e'].get('ms_per_step'))"
done
