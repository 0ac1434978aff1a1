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
    T* g_controls;       // (B,T,2)
    T* g_joint_angles;   // (B,T,4) moving-flipper variant only
    T* g_x0;             // (B,3)
    T* g_xd0;            // (B,3)
    T* g_R0;             // (B,9)
    T* g_om0;            // (B,3)
};

}  // namespace mfb
