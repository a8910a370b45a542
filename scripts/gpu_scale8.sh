#!/bin/bash
# N = 4 and 8 scaling lines (default layout only)
mkdir -p gpurun_out
for n in 4 8; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; cut -c1-330 gpurun_out/scale_n$n.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/scale_n$n.err | tail -3
done
