#!/bin/bash
# Run on the GPU box (under gpurun): bench line, ncu launch list of the same command, one full capture per hot kernel.
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tail -1 > gpurun_out/bench.json
cat gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_ -s 4 -c 2 -o gpurun_out/prof_rollout \
    python tools/profile_target.py 4096 3 > gpurun_out/prof.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.csv
