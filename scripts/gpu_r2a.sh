#!/bin/bash
# round 2, first visit: certificate tests, the native-engine parity tests, bench with and without the certificate
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_cert.py tests/test_gpu_native.py tests/test_gpu_parity.py -x -q > gpurun_out/pytest_cert.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cert.log
tail -15 gpurun_out/pytest_cert.log
for c in 1 0; do
EMF_RAY_CERT=$c timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cert$c.json 2> gpurun_out/bench_cert$c.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_cert$c.json").readline())
print("cert=$c", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stages_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4))
PY
tail -2 gpurun_out/bench_cert$c.err
done
