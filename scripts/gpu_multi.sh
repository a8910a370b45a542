#!/bin/bash
# One multi-GPU box visit: the 2-rank parity tests, then bench.py at N = 1..$1 (default 2), both background layouts.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm --format=csv > gpurun_out/smi_multi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -4 gpurun_out/pytest_multi.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; cut -c1-400 gpurun_out/scale_n1.json
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; cut -c1-400 gpurun_out/scale_n$n.json; tail -3 gpurun_out/scale_n$n.err
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $n --steps 30 --warmup 5 --background-on-rank0 > gpurun_out/scale_bg0_n$n.json 2> gpurun_out/scale_bg0_n$n.err; cut -c1-400 gpurun_out/scale_bg0_n$n.json; tail -3 gpurun_out/scale_bg0_n$n.err
  fi
done
