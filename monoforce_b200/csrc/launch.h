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
// zeroes g.g_scratch, runs the adjoint kernel (single-sweep when the contact_sum tape is present and the geometry
// is static, else the three-pass kernel) and adds the map-gradient scratch into g.g_z / g.g_mu
template <typename T, int VARIANT>
LaunchError launch_rollout_bwd(const RolloutArgs<T>& args, const AdjointArgs<T>& g, cudaStream_t stream);

void count_launch();

}  // namespace mfb
