#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:'k_ray|k_integrate' -s 18 -c 6 --csv --log-file gpurun_out/ray_cert.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ray.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open("gpurun_out/ray_cert.csv") if not l.startswith("==")))
h=None
for r in rows:
    if len(r)>5 and r[0]=="ID": h=r; continue
    if h and len(r)==len(h):
        d=dict(zip(h,r)); print(d["ID"], d["Kernel Name"][:40], d["Metric Name"], d["Metric Value"])
PY

