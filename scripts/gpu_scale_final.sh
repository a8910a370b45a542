#!/bin/bash
# final scaling lines at N = 8 and 4 (default configuration: replicated background, NVLink peer-memory exchanges)
mkdir -p gpurun_out
for n in 8 4; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2963$n bench.py --gpus $n --steps 40 --warmup 5 > gpurun_out/scale_final_n$n.json 2> gpurun_out/scale_final_n$n.err
  wc -l gpurun_out/scale_final_n$n.json; python -c "
import json; d=json.loads(open('gpurun_out/scale_final_n$n.json').readline())
print('N=$n', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['gpu_launches'], d['config']['parallelism'][-40:])"
  grep -v "OMP_NUM\|\*\*\*" gpurun_out/scale_final_n$n.err | tail -2
done
