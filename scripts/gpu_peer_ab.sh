#!/bin/bash
# multi-GPU A/B: NVLink peer-memory exchange vs NCCL collectives at N ranks: scripts/gpu_peer_ab.sh N
N=${1:-2}
mkdir -p gpurun_out
for mode in peer nccl peer nccl; do
  extra=""; [ $mode = nccl ] && extra="--nccl-exchange"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 40 --warmup 5 $extra 2>gpurun_out/peer_ab.err | grep "^{" > gpurun_out/scale_${mode}_n$N.json
  python -c "
import json; d=json.loads(open('gpurun_out/scale_${mode}_n$N.json').readline())
print('$mode N=$N', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'host', round(d['host_issue_ms_per_step'],3), d['config']['parallelism'][-40:])"
  grep -v "OMP_NUM\|\*\*\*" gpurun_out/peer_ab.err | tail -2
done
