// Instantiations + PPL dispatcher of the forward rollout kernel (K1) for one
// (scalar type, integrator variant) pair: -DMFB_INST_T=... -DMFB_INST_VARIANT=...
#include "launch.h"
#include "rollout_fwd.cuh"

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
    const size_t smem = forces ? RowStage<T>::smem_bytes(a.N, kFwdWarps) : 0;
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

template <>
LaunchError launch_rollout_fwd<MFB_INST_T, MFB_INST_VARIANT>(const RolloutArgs<MFB_INST_T>& a, cudaStream_t st) {
    using T = MFB_INST_T;
    constexpr int V = MFB_INST_VARIANT;
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
