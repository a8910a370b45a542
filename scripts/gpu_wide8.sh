#!/bin/bash
N=${1:-8}
run() {  # tag env
  tag=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$tag.json").readline())
    print("$tag", round(d["ms_per_step"],4), {k:(round(v,4) if isinstance(v,float) else None) for k,v in d["stages_ms"].items() if k!="note"}, "e2e", round(d["e2e"]["ms_per_step"],4), "parity_ok", (d.get("parity_vs_n1") or {}).get("ok"), "errs", d.get("exchange_errors"))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/scale_$tag.err").read()[-800:])
PY
}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_cert.py -x -q -k "wide or schedule" 2>&1 | tail -2
run n${N}_wide EMF_RAY_WIDE=1
