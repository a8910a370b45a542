#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cert.py tests/test_gpu_native.py -x -q > gpurun_out/pytest_cert.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cert.log
tail -8 gpurun_out/pytest_cert.log
timeout 300 python scripts/diag_cert.py 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_ray' -s 18 -c 3 --csv --log-file gpurun_out/ray_cert.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ray.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open("gpurun_out/ray_cert.csv") if not l.startswith("==")))
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print(d["ID"], d["Kernel Name"][:30], d["Metric Name"], d["Metric Value"])
PY
for c in 1 0; do
EMF_RAY_CERT=$c timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cert$c.json 2> gpurun_out/bench_cert$c.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_cert$c.json").readline())
print("cert=$c", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4))
PY
done
