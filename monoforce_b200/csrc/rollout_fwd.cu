// Instantiations + PPL dispatcher of the forward rollout kernel (K1) for one
// (scalar type, integrator variant) pair: -DMFB_INST_T=... -DMFB_INST_VARIANT=...
#include "launch.h"
#include <cstdlib>
#include "rollout_fwd.cuh"
#include "rollout_fwd_wide.cuh"

#ifndef MFB_INST_T
#error "compile with -DMFB_INST_T=float|double -DMFB_INST_VARIANT=0|1"
#endif

namespace mfb {

template <typename T, int PPL, int VARIANT>
static LaunchError launch_ppl(const RolloutArgs<T>& a, cudaStream_t st) {
    const dim3 grid((a.B + kFwdWarps - 1) / kFwdWarps), block(kFwdWarps * 32);
    const bool forces = a.Fs != nullptr, cost = a.cost != nullptr;
    if (forces && ((((uintptr_t)a.Fs) | ((uintptr_t)a.Ff)) & 15))
        return {"F_springs / F_frictions must be 16-byte aligned (rows leave the SM as TMA bulk stores)"};
    const size_t smem = forces ? RowStage<T>::smem_bytes(a.N, kFwdWarps, VARIANT) : 0;
    auto go = [&](auto kern) -> LaunchError {
        // static (point table) + dynamic (row images) can exceed the 48 KB default: always opt in
        if (smem > 0 &&
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return {"cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed"};
        kern<<<grid, block, smem, st>>>(a);
        count_launch();
        return {nullptr};
    };
    if (a.joint_angles) {
        // moving flippers (marv): one instantiation per integrator, forces + cost handled at run time would bloat the
        // build, so this variant always materialises forces and never fuses the cost
        if (!forces || cost) return {"the moving-flipper variant needs F_springs/F_frictions and does not fuse the cost"};
        return go(rollout_fwd_kernel<T, PPL, VARIANT, true, false, true>);
    }
    if (VARIANT == kOdeintEuler) {
        if (cost) return {"cost output is defined for the step-loop variant only"};
        if (forces) return go(rollout_fwd_kernel<T, PPL, VARIANT, true, false>);
        return go(rollout_fwd_kernel<T, PPL, VARIANT, false, false>);
    }
    if (forces && cost) return go(rollout_fwd_kernel<T, PPL, VARIANT, true, true>);
    if (forces) return go(rollout_fwd_kernel<T, PPL, VARIANT, true, false>);
    if (cost) return go(rollout_fwd_kernel<T, PPL, VARIANT, false, true>);
    return go(rollout_fwd_kernel<T, PPL, VARIANT, false, false>);
}

// Small batches: one CTA per trajectory, two contact points per thread (K1w).  Below ~2 warps per scheduler the one-warp-per-
// trajectory kernel is bound by single-warp latency; K1w shortens the instruction stream each warp walks per step.  Measured
// crossover on B200 (tools/fwd_crossover.py, T = 500, marv): step loop B ~ 512 (1.05 vs 1.09 ms; 0.69 vs 1.08 ms at B = 64),
// odeint variant B ~ 2048 (0.65 vs 1.26 ms at B = 64).  MFB_FWD_WIDE_MAX_B overrides the batch-size threshold (0 disables the
// wide kernel; tests force either path).
static int wide_max_b(int variant) {
    const char* e = getenv("MFB_FWD_WIDE_MAX_B");
    return e ? atoi(e) : (variant == kOdeintEuler ? 1024 : 512);
}

template <typename T, int VARIANT>
static LaunchError launch_wide(const RolloutArgs<T>& a, cudaStream_t st) {
    const dim3 grid(a.B), block(((a.N + 32 * kWidePts - 1) / (32 * kWidePts)) * 32);
    const bool forces = a.Fs != nullptr, cost = a.cost != nullptr;
    auto go = [&](auto kern) -> LaunchError {
        kern<<<grid, block, 0, st>>>(a);
        count_launch();
        return {nullptr};
    };
    if (VARIANT == kOdeintEuler) {
        if (cost) return {"cost output is defined for the step-loop variant only"};
        if (forces) return go(rollout_fwd_wide_kernel<T, VARIANT, true, false>);
        return go(rollout_fwd_wide_kernel<T, VARIANT, false, false>);
    }
    if (forces && cost) return go(rollout_fwd_wide_kernel<T, VARIANT, true, true>);
    if (forces) return go(rollout_fwd_wide_kernel<T, VARIANT, true, false>);
    if (cost) return go(rollout_fwd_wide_kernel<T, VARIANT, false, true>);
    return go(rollout_fwd_wide_kernel<T, VARIANT, false, false>);
}

template <>
LaunchError launch_rollout_fwd<MFB_INST_T, MFB_INST_VARIANT>(const RolloutArgs<MFB_INST_T>& a, cudaStream_t st) {
    using T = MFB_INST_T;
    constexpr int V = MFB_INST_VARIANT;
    if (!a.joint_angles && a.B <= wide_max_b(V)) return launch_wide<T, V>(a, st);
    const int ppl = (a.N + 31) / 32;
    switch (ppl) {
        case 1: return launch_ppl<T, 1, V>(a, st);
        case 2: return launch_ppl<T, 2, V>(a, st);
        case 3: return launch_ppl<T, 3, V>(a, st);
        case 4: return launch_ppl<T, 4, V>(a, st);
        case 5: return launch_ppl<T, 5, V>(a, st);
        case 6: return launch_ppl<T, 6, V>(a, st);
        case 7: return launch_ppl<T, 7, V>(a, st);
        case 8: return launch_ppl<T, 8, V>(a, st);
        default: return {"number of contact points must be in [1, 256]"};
    }
}

}  // namespace mfb
