// Instantiations + PPL dispatcher of the adjoint rollout kernel (K2) for one
// (scalar type, integrator variant) pair: -DMFB_INST_T=... -DMFB_INST_VARIANT=...
#include "launch.h"
#include <cstdlib>
#include "rollout_bwd_sweep.cuh"

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

// single-sweep adjoint (K2s): static geometry + the forward's contact_sum tape.  Up to MFB_BWD_WIDE_MAX_B trajectories (default
// 512) the four warps of a CTA share one trajectory (SPLIT instantiation): small training batches are bound by the latency of
// one warp, not by throughput.  Measured on B200 (tools/bwd_crossover.py, T = 500, marv): 2.4 -> 1.3 ms up to B = 128,
// 2.49 vs 2.23 ms at B = 512, 2.99 vs 3.97 ms at B = 1024.
static int bwd_wide_max_b() {
    const char* e = getenv("MFB_BWD_WIDE_MAX_B");
    return e ? atoi(e) : 512;
}

template <typename T, int VARIANT>
static LaunchError launch_sweep(const RolloutArgs<T>& a, const AdjointArgs<T>& g, cudaStream_t st) {
    const bool wide = a.B <= bwd_wide_max_b();
    const dim3 grid(wide ? a.B : (a.B + kSweepWarps - 1) / kSweepWarps), block(kSweepWarps * 32);
    const int slots = ((a.N + 31) / 32) * 32;
    const size_t smem = g.g_cells ? (size_t)(wide ? 1 : kSweepWarps) * 3 * slots * sizeof(Quad<T>) : 0;
    auto go = [&](auto kern) -> LaunchError {
        if (smem > 0 &&
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return {"cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed"};
        kern<<<grid, block, smem, st>>>(a, g);
        count_launch();
        return {nullptr};
    };
    if constexpr (VARIANT == kStepLoop) {
        if (g.g_Fs || g.g_Ff)
            return wide ? go(rollout_bwd_sweep_kernel<T, VARIANT, true, kSweepWarps>) : go(rollout_bwd_sweep_kernel<T, VARIANT, true, 1>);
    }
    return wide ? go(rollout_bwd_sweep_kernel<T, VARIANT, false, kSweepWarps>) : go(rollout_bwd_sweep_kernel<T, VARIANT, false, 1>);
}

template <>
LaunchError launch_rollout_bwd<MFB_INST_T, MFB_INST_VARIANT>(const RolloutArgs<MFB_INST_T>& a,
                                                            const AdjointArgs<MFB_INST_T>& g_in, cudaStream_t st) {
    using T = MFB_INST_T;
    constexpr int V = MFB_INST_VARIANT;
    AdjointArgs<T> g = g_in;
    const long long HW = (long long)a.H * a.W;
    const bool shared_map = a.map_stride == 0;
    const bool sweep = a.Csum != nullptr && a.joint_angles == nullptr && !(V == kOdeintEuler && (g.g_Fs || g.g_Ff));
    const int per_cell = sweep ? kGradRec : 2;
    g.g_maps = g.g_cells = nullptr;
    if (g.g_scratch) {
        if (cudaMemsetAsync(g.g_scratch, 0, (size_t)(g.n_maps * HW * per_cell) * sizeof(T), st) != cudaSuccess)
            return {"cudaMemsetAsync(map-gradient scratch) failed"};
        if (sweep) { g.g_cells = g.g_scratch; g.g_cells_stride = shared_map ? 0 : HW * kGradRec; }
        else { g.g_maps = g.g_scratch; g.g_maps_stride = shared_map ? 0 : HW * 2; }
    }
    LaunchError e{nullptr};
    if (sweep) {
        e = launch_sweep<T, V>(a, g, st);
    } else {
        const int ppl = (a.N + 31) / 32;
        switch (ppl) {
            case 1: e = launch_ppl<T, 1, V>(a, g, st); break;
            case 2: e = launch_ppl<T, 2, V>(a, g, st); break;
            case 3: e = launch_ppl<T, 3, V>(a, g, st); break;
            case 4: e = launch_ppl<T, 4, V>(a, g, st); break;
            case 5: e = launch_ppl<T, 5, V>(a, g, st); break;
            case 6: e = launch_ppl<T, 6, V>(a, g, st); break;
            case 7: e = launch_ppl<T, 7, V>(a, g, st); break;
            case 8: e = launch_ppl<T, 8, V>(a, g, st); break;
            default: return {"number of contact points must be in [1, 256]"};
        }
    }
    if (e.msg) return e;
    if (g.g_scratch) {
        const long long n = g.n_maps * HW;
        const int block = 256;
        const int grid = (int)((n + block - 1) / block < 148 * 16 ? (n + block - 1) / block : 148 * 16);
        if (sweep) finalize_map_grads_kernel<T><<<grid, block, 0, st>>>(g.g_scratch, g.g_z, g.g_mu, (int)g.n_maps, a.H, a.W);
        else scatter_map_grads_kernel<T><<<grid, block, 0, st>>>(g.g_scratch, g.g_z, g.g_mu, n);
        count_launch();
    }
    return {nullptr};
}

}  // namespace mfb
