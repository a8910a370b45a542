#!/bin/bash
# scaling lines on an N-GPU box: scripts/gpu_scale.sh N [tests] [ab]   (writes gpurun_out/scale_n<N>.json; "ab": also EMF_RAY_SCHED=0)
N=${1:-2}
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then
  timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_n$N.log
  tail -4 gpurun_out/pytest_multi_n$N.log
fi
run() {  # tag env
  tag=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$tag.json").readline())
    print("$tag", round(d["ms_per_step"],4), {k:(round(v,4) if isinstance(v,float) else None) for k,v in d["stages_ms"].items() if k!="note"}, "e2e", round(d["e2e"]["ms_per_step"],4), "parity_ok", (d.get("parity_vs_n1") or {}).get("ok"), "errs", d.get("exchange_errors"), "clocks", d.get("clocks"))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/scale_$tag.err").read()[-800:])
PY
}
run n$N EMF_DUMMY=1
if [ "$2" = "ab" ] || [ "$3" = "ab" ]; then run n${N}_nosched EMF_RAY_SCHED=0; fi
