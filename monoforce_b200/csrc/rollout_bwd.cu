// Instantiations + PPL dispatcher of the adjoint rollout kernel (K2) for one
// (scalar type, integrator variant) pair: -DMFB_INST_T=... -DMFB_INST_VARIANT=...
#include "launch.h"
#include "rollout_bwd.cuh"

#ifndef MFB_INST_T
#error "compile with -DMFB_INST_T=float|double -DMFB_INST_VARIANT=0|1"
#endif

namespace mfb {

template <typename T, int PPL, int VARIANT>
static LaunchError launch_ppl(const RolloutArgs<T>& a, const AdjointArgs<T>& g, cudaStream_t st) {
    const dim3 grid((a.B + kBwdWarps - 1) / kBwdWarps), block(kBwdWarps * 32);
    const size_t smem = g.g_maps ? kBwdWarps * sizeof(MapGradCache<T>) : 0;
    auto go = [&](auto kern) -> LaunchError {
        if (smem > 0 &&
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return {"cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed"};
        kern<<<grid, block, smem, st>>>(a, g);
        count_launch();
        return {nullptr};
    };
    if (a.joint_angles) return go(rollout_bwd_kernel<T, PPL, VARIANT, true, true>);     // moving flippers
    if (g.g_Fs || g.g_Ff) return go(rollout_bwd_kernel<T, PPL, VARIANT, true>);
    return go(rollout_bwd_kernel<T, PPL, VARIANT, false>);
}

template <>
LaunchError launch_rollout_bwd<MFB_INST_T, MFB_INST_VARIANT>(const RolloutArgs<MFB_INST_T>& a,
                                                            const AdjointArgs<MFB_INST_T>& g, cudaStream_t st) {
    using T = MFB_INST_T;
    constexpr int V = MFB_INST_VARIANT;
    const int ppl = (a.N + 31) / 32;
    switch (ppl) {
        case 1: return launch_ppl<T, 1, V>(a, g, st);
        case 2: return launch_ppl<T, 2, V>(a, g, st);
        case 3: return launch_ppl<T, 3, V>(a, g, st);
        case 4: return launch_ppl<T, 4, V>(a, g, st);
        case 5: return launch_ppl<T, 5, V>(a, g, st);
        case 6: return launch_ppl<T, 6, V>(a, g, st);
        case 7: return launch_ppl<T, 7, V>(a, g, st);
        case 8: return launch_ppl<T, 8, V>(a, g, st);
        default: return {"number of contact points must be in [1, 256]"};
    }
}

template <>
void launch_scatter_map_grads<MFB_INST_T, MFB_INST_VARIANT>(const MFB_INST_T* g2, MFB_INST_T* g_z, MFB_INST_T* g_mu,
                                                            long long n, cudaStream_t st) {
    const int block = 256;
    const int grid = (int)((n + block - 1) / block < 148 * 16 ? (n + block - 1) / block : 148 * 16);
    scatter_map_grads_kernel<MFB_INST_T><<<grid, block, 0, st>>>(g2, g_z, g_mu, n);
    count_launch();
}

}  // namespace mfb
