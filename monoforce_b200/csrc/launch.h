// Host-side launchers implemented in rollout_fwd.cu / rollout_bwd.cu, used by c_api.cu.
// Each (scalar type, variant) pair is compiled as its own translation unit
// (-DMFB_INST_T=float|double -DMFB_INST_VARIANT=0|1) so the build parallelises.
#pragma once
#include "rollout_common.cuh"

namespace mfb {

struct LaunchError {
    const char* msg;   // nullptr on success
};

template <typename T, int VARIANT>
LaunchError launch_rollout_fwd(const RolloutArgs<T>& args, cudaStream_t stream);

template <typename T> struct AdjointArgs;
template <typename T, int VARIANT>
LaunchError launch_rollout_bwd(const RolloutArgs<T>& args, const AdjointArgs<T>& g, cudaStream_t stream);

// adds the interleaved (z, mu) gradient scratch into the caller's maps (instantiated next to the kernels)
template <typename T, int VARIANT>
void launch_scatter_map_grads(const T* g2, T* g_z, T* g_mu, long long n, cudaStream_t stream);

void count_launch();

}  // namespace mfb
