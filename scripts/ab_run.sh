#!/bin/bash
# run bench.py once per A/B library: scripts/ab_run.sh tag1 tag2 ...   (prints ms/step and the stage times)
for t in "$@"; do
  EMF_B200_LIB=$PWD/gpurun_ab/libemf_$t.so timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$t', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stages_ms'].items() if isinstance(v,(int,float))}, 'e2e', d['e2e'].get('ms_per_step'))"
done
