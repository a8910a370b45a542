#!/bin/bash
# One GPU-box visit (round 2): parity tests, both bench arms (+ the unchanged-caller configuration), the ncu launch list and one
# full capture of the frame's kernels.  scripts/summarise_profiles.py r2 turns gpurun_out/ into profiles/r2_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 python bench.py --impl unchanged-caller --steps 20 --warmup 3 > gpurun_out/bench_unchanged.json 2> gpurun_out/bench_unchanged.err; cut -c1-200 gpurun_out/bench_unchanged.json
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; cut -c1-400 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
if [ "$1" = "prof" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'emfb|k_' -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_integrate|k_brick|k_raycast|k_assoc|k_composite|k_depth' -s 24 -c 6 -f -o gpurun_out/prof_r2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
for c in 2 3 5; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err; cut -c1-160 gpurun_out/bench_cfg$c.json; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python scripts/diag_stats.py > gpurun_out/diag_stats.log 2>&1; grep "raycast bg:\|raycast objs:" gpurun_out/diag_stats.log
fi
ls gpurun_out | head -50
