#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "n_rank" > gpurun_out/pytest_multi8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi8.log
tail -4 gpurun_out/pytest_multi8.log
run() {  # n tag extra-args
  n=$1; tag=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$tag.json").readline())
    print("$tag", round(d["ms_per_step"],4), {k:(round(v,4) if isinstance(v,float) else None) for k,v in d["stages_ms"].items() if k!="note"}, "e2e", round(d["e2e"]["ms_per_step"],4), "parity_ok", (d.get("parity_vs_n1") or {}).get("ok"), "errs", d.get("exchange_errors"))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/scale_$tag.err").read()[-800:])
PY
}
run 8 n8_cert
run 8 n8_plain --ray-certificate 0
run 4 n4_cert
run 4 n4_plain --ray-certificate 0
run 2 n2_plain --ray-certificate 0
