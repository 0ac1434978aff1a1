#!/bin/bash
# Round-2 evidence, run on the GPU box (under gpurun): ncu captures of the depthwise / heads / rollout kernels, launch lists with DRAM
# bytes for both encoder sizes, the forward-kernel crossover and the depthwise layer table.  tools/summarize_*.py turn them into profiles/.
set -x
for L in 9 2 1; do timeout 300 ncu --set full --import-source on --clock-control none -k regex:dwconv --launch-skip 3 -c 1 -o gpurun_out/prof_dwf$L -f python tools/dw_bench.py --cfg4 --only $L --reps 3 > gpurun_out/prof_dwf$L.log 2>&1; done
python tools/dw_bench.py --cfg4 > gpurun_out/dw_bench_cfg4.txt; python tools/dw_bench.py > gpurun_out/dw_bench_default.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_bn_act_kernel -s 103 -c 1 -o gpurun_out/prof_heads -f python tools/encoder_once.py --cfg4 1 > gpurun_out/prof_heads.log 2>&1
for c in default cfg4; do f=""; [ $c = cfg4 ] && f="--cfg4"; timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/enc_launches_$c.csv python tools/encoder_once.py $f > gpurun_out/once_$c.log 2>&1; done
timeout 300 python tools/fwd_crossover.py --out gpurun_out/fwd_crossover.json > gpurun_out/fwd_crossover.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rollout_ -s 4 -c 2 -o gpurun_out/prof_rollout -f python tools/profile_target.py 4096 3 > gpurun_out/prof.log 2>&1
echo collected
