#!/bin/bash
# multi-GPU visit: scripts/gpu_multi2.sh N  (tests at N ranks, then the scaling bench lines)
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -6 gpurun_out/pytest_multi.log
for n in $(seq 1 $N); do
  case $n in 1|2|4|8) ;; *) continue;; esac
  if [ $n = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n"; fi
  timeout 600 $cmd --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_n$n.json").readline())
    print("N=$n", round(d["ms_per_step"],4), d["stages_ms"] and {k:(round(v,4) if isinstance(v,float) else None) for k,v in d["stages_ms"].items() if k!="note"}, "e2e", round(d["e2e"]["ms_per_step"],4), "parity", d.get("parity_vs_n1"), "errs", d.get("exchange_errors"))
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/scale_n$n.err").read()[-1500:])
PY
done
