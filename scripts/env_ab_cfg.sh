#!/bin/bash
# bench.py --config C once per environment setting: scripts/env_ab_cfg.sh C "EMF_RAY_SCHED=0" ...
c=$1; shift
for t in "$@"; do
  env $t timeout 300 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('cfg$c $t', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stages_ms'].items() if isinstance(v,(int,float))}, 'e2e', d['e2e'].get('ms_per_step'))"
done
