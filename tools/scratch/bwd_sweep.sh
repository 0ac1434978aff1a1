#!/bin/bash
for v in "" bwd4; do
  if [ -z "$v" ]; then unset MFB_LIB_PATH; echo "== default (bwd 3 CTAs/SM, 168 regs)"; else export MFB_LIB_PATH=$PWD/tools/scratch/libmfb_$v.so; echo "== $v (bwd 4 CTAs/SM, 128 regs, spills)"; fi
  python tools/quick_time.py 2>&1 | grep -E "grads to z and friction|controls only"
done
unset MFB_LIB_PATH
ncu --set full --clock-control none --import-source on -k regex:rollout_ -s 4 -c 2 -o gpurun_out/prof_rollout3 python tools/profile_target.py 4096 3 > gpurun_out/prof.log 2>&1; tail -1 gpurun_out/prof.log
