// Argument block of the adjoint kernel (K2); see include/monoforce_b200.h mfb_rollout_grads.
#pragma once
#include "rollout_common.cuh"

namespace mfb {

template <typename T>
struct AdjointArgs {
    // incoming gradients (nullptr == zero)
    const T* g_Xs;       // (B,T,3)
    const T* g_Xds;      // (B,T,3)
    const T* g_Rs;       // (B,T,9)
    const T* g_Oms;      // (B,T,3)
    const T* g_Fs;       // (B,T,N,3)
    const T* g_Ff;       // (B,T,N,3)
    const T* g_x0z;      // (B,)
    // outgoing gradients (nullptr == not wanted)
    T* g_maps;           // (B|1,H,W,2) zero-initialised scratch: (d/dz, d/dfriction) interleaved per cell, accumulated
                         // with vector atomics; scatter_map_grads_kernel adds it into the caller's g_z_grid / g_friction.
                         // nullptr == no map gradient wanted
    long long g_maps_stride;   // elements between two trajectories' scratch maps (0 = shared)
    // single-sweep kernel: the same zero-initialised scratch viewed as (B|1,H,W,8) per-cell corner records
    // (z00 z10 z01 z11 | m00 m10 m01 m11), finalize_map_grads_kernel adds them into g_z / g_mu
    T* g_cells;
    long long g_cells_stride;
    T* g_z;              // (B|1,H,W) caller's d/dz_grid (accumulated) or nullptr
    T* g_mu;             // (B|1,H,W) caller's d/dfriction (accumulated) or nullptr
    long long g_dir_stride;    // elements between two trajectories' maps in g_z / g_mu (0 = shared)
    T* g_scratch;        // zero-initialised by the launcher: max(2, 8) * n_maps * H * W scalars, or nullptr (no map gradient)
    long long n_maps;
    T* g_controls;       // (B,T,2)
    T* g_joint_angles;   // (B,T,4) moving-flipper variant only
    T* g_x0;             // (B,3)
    T* g_xd0;            // (B,3)
    T* g_R0;             // (B,9)
    T* g_om0;            // (B,3)
};

}  // namespace mfb
