#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_accel.py tests/test_gpu_parity.py tests/test_gpu_native.py tests/test_gpu_frame.py tests/test_gpu_vs_reference.py -x -q > gpurun_out/pytest_bricks.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_bricks.log
tail -8 gpurun_out/pytest_bricks.log
for c in 1 0; do
EMF_INT_BRICKS=$c timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_bricks$c.json 2> gpurun_out/bench_bricks$c.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_bricks$c.json").readline())
print("bricks=$c", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4), d["roofline"].get("counters"))
PY
tail -2 gpurun_out/bench_bricks$c.err
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_integrate|k_brick|k_depth' -s 18 -c 3 --csv --log-file gpurun_out/int_bricks.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_int.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open("gpurun_out/int_bricks.csv") if not l.startswith("==")))
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print(d["ID"], d["Kernel Name"][:30], d["Metric Name"], d["Metric Value"])
PY
