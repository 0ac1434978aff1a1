#!/bin/bash
# profiles/r02_* from the files tools/collect_profiles_r02.sh left in gpurun_out/ (run here, no GPU needed)
set -e
cd "$(dirname "$0")/.."
python tools/summarize_encoder_profiles.py r02 > /dev/null
cp profiles/r02_bench_line.json /tmp/r02_bench_line.keep            # summarize_profiles.py rewrites it from gpurun_out/bench.json
python tools/summarize_profiles.py r02 > /dev/null
cp /tmp/r02_bench_line.keep profiles/r02_bench_line.json
{
echo; echo "## K7d, TMA-fed depthwise kernel (csrc/dwconv_tma.cu): ncu --set full of three EfficientNet-B0 layers at BASELINE config 4 (64 images)"; echo
echo 'Command: `ncu --set full --import-source on --clock-control none -k regex:dwconv --launch-skip 3 -c 1 python tools/dw_bench.py --cfg4 --only <layer> --reps 3`'; echo
for L in 9 2 1; do case $L in 9) t="layer 9: 672 channels, 5x5, stride 1, 32x32 (64-channel slabs, two 8-row passes per TMA tile)";; 2) t="layer 2: 144 channels, 3x3, stride 1, 128x128";; 1) t="layer 1: 96 channels, 3x3, stride 2, 256x256 -> 128x128 (load-bound)";; esac; echo "### $t"; echo; echo '```'; python tools/ncu_brief.py gpurun_out/prof_dwf$L.ncu-rep 8 2>&1 | head -30; echo '```'; echo; done
echo "## Depthwise layers one by one (tools/dw_bench.py, CUDA events, L2 flushed between runs)"; echo; echo '```'; echo "# BASELINE config 4 (64 images 512x512)"; cat gpurun_out/dw_bench_cfg4.txt; echo "# lss_cfg.yaml (64 images 256x416)"; cat gpurun_out/dw_bench_default.txt; echo '```'
echo; echo "## K4 on a memory-bound layer (MBConv block-1 expand, pixel-folded x4: 64 x 256 x 64 x 64 -> 384, SiLU): where the epilogue's issue slots went"; echo
echo 'Command: `ncu --set full --import-source on --clock-control none -k regex:conv_bn_act --launch-skip 5 -c 1 python tools/conv_experiment.py "b1_expand folded"`; SASS opcode counts from `ncu --page source --csv` divided by the 12.58 M (32 pixels x 1 channel) output units.'; echo
echo "Before the epilogue pass (commit 'K4 header comments'): 295.9 us, 194.2 M warp-instructions = 15.4 per unit, issue slots 66 % busy:"; echo '```'
echo "IMAD 2.01  FFMA 2.00  BRA 1.75  SYNCS 1.36  YIELD 1.34  MUFU 1.06  VIADD 1.05  FADD 1.00  FMUL 1.00  IADD3 0.90  ISETP 0.69  LDS 0.63  LOP3 0.58  R2UR 0.56  PRMT 0.50  F2FP 0.50"
echo '```'
echo "BRA / SYNCS / YIELD = the polling loops of the TMA producer and the MMA issuer; 1.0 VIADD = fallback values of a development switch; 1.0 FFMA + 1.0 IMAD (both predicated off) = the residual add in a layer without residual.  After (relaxed polling, residual as a template parameter, switch removed, SiLU on pre-halved parameters, packed f32x2, per-tile store indices): 232 us = 4.0 TB/s (CUDA events, tools/conv_experiment.py)."
} >> profiles/r02_ncu_encoder_summary.md
{
echo; echo "## Small batches: K1 vs K1w at the planner's size (64 trajectories x 500 steps, shared 128x128 map, step loop, forces materialised)"; echo
echo 'Command: `[MFB_FWD_WIDE_MAX_B=0] ncu --set full --import-source on --clock-control none -k regex:rollout_fwd --launch-skip 2 -c 1 python tools/profile_small.py`'; echo
echo "### K1 (one warp per trajectory; 16 CTAs)"; echo; echo '```'; python tools/ncu_brief.py gpurun_out/small_warp.ncu-rep 6 2>&1 | head -26; echo '```'; echo
echo "### K1w (one CTA per trajectory, two contact points per thread; 64 CTAs x 128 threads)"; echo; echo '```'; python tools/ncu_brief.py gpurun_out/small_wide2.ncu-rep 6 2>&1 | head -26; echo '```'; echo
echo "### Forward-only latency per batch size, both kernels (tools/fwd_crossover.py; wall clock of the DPhysics call incl. the cell table)"; echo; echo '```'; cat gpurun_out/fwd_crossover.txt; echo '```'
echo; echo "## Small batches: single-sweep adjoint, one warp vs one CTA per trajectory (tools/bwd_crossover.py; CUDA events around loss.backward(), T = 500, marv, shared 128x128 map)"; echo; echo '```'; cat gpurun_out/bwd_crossover.txt; echo '```'
cat <<'EOF'

## Instruction audit of the two headline kernels (ncu source page of `prof_rollout.ncu-rep`, per trajectory-step)

The K4 epilogue turned out to spend ~20 % of its issue slots on instructions that did nothing (predicated-off residual code, a
development switch, polling loops; see the encoder summary).  The same audit on the rollout kernels finds nothing of the kind:

| kernel | warp-instructions / step | FFMA + FMUL + FADD | fully predicated-off |
|---|---|---|---|
| `rollout_bwd_sweep_kernel<float,0,0>` (K2s) | 3174 | 1952 (61 %) | 46 (1.4 %: BRA 11, IMAD 11, LDG 6, FADD 6, ...) |
| `rollout_fwd_kernel<float,7,0,1,1,0>` (K1) | 1265 | 669 (53 %) | 2 |

K2s top opcodes: FFMA 1053, FMUL 627, FADD 272, FSEL 111, LDS 107, ISETP 101, IMAD 88, FMNMX 86, BRA 80, LDCU 72, FSETP 60, SHFL 55, LDG 51.
K1 top opcodes: FFMA 336, FMUL 201, FADD 132, FMNMX 93, IADD3 49, IMAD 43, STS 42, MUFU 26, ISETP 24, SEL 24, LDCU 24, LDG 23, SHFL 22, STG 22.
EOF
} >> profiles/r02_ncu_rollout_summary.md
python tools/launch_bw.py gpurun_out/enc_launches_cfg4.csv > profiles/r02_encoder_launch_bw_cfg4.txt
python tools/launch_bw.py gpurun_out/enc_launches_default.csv > profiles/r02_encoder_launch_bw_default.txt
cp gpurun_out/fwd_crossover.json profiles/r02_fwd_crossover.json
cp gpurun_out/bwd_crossover.json profiles/r02_bwd_crossover.json
echo "profiles refreshed"
