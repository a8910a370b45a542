#!/bin/bash
# full ncu capture of the frame's kernels (one steady-state frame): scripts/gpu_prof.sh <tag> [kernel regex]
tag=${1:-r2}; rx=${2:-k_ray|k_integrate|k_brick}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s ${3:-18} -c ${4:-3} -f -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
tail -3 gpurun_out/ncu_full_$tag.log | cut -c1-300
ls -la gpurun_out/prof_$tag.ncu-rep
