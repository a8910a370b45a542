#!/bin/bash
mkdir -p gpurun_out
run() { # mode N
  extra=""; [ $1 = nccl ] && extra="--nccl-exchange"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $2 --steps 40 --warmup 5 $extra 2>gpurun_out/peer_ab.err | grep "^{" > gpurun_out/scale_$1_n$2.json
  python -c "
import json; d=json.loads(open('gpurun_out/scale_$1_n$2.json').readline())
print('$1 N=$2', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'host', round(d['host_issue_ms_per_step'],3), d['config']['parallelism'][-40:])"
  grep -v "OMP_NUM\|\*\*\*" gpurun_out/peer_ab.err | tail -2
}
run peer 8; run nccl 8; run peer 4; run peer 8
