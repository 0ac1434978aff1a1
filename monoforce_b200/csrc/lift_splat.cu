// K5: fused "lift" + "splat" of the Lift-Splat-Shoot terrain encoder.
//
// Reference (monoforce/src/monoforce/models/terrain_encoder/):
//   lift   lss.py:63-71    depth = softmax(logits[:, :D]);  lifted = depth (x) feats  -> (BN, C, D, fH, fW)
//   splat  lss.py:238-280  voxel index per frustum point, drop out-of-grid points, sort by voxel rank,
//                          segment-sum with the cumsum trick (utils.py:144-181), scatter into (B, C, Z, X, Y)
// Here: one warp per frustum pixel (b, n, h, w).  The D depth logits are soft-maxed in registers (lane d % 32
// owns logit d), lane l owns channels 2l and 2l+1 of the C = 64 features, and for every depth bin whose point
// falls inside the grid the warp adds depth_d * feats to the channels-last BEV cell with ONE coalesced 256-byte
// vector reduction (red.global.add.v2.f32).  The lifted tensor and the argsort never exist.
// Backward is the matching gather: no atomics.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/monoforce_b200.h"

namespace mfb {
void count_launch();
int fail_status(int code, const std::string& msg);

constexpr int kMaxDepthSlots = 4;     // D <= 128 depth bins
constexpr unsigned kAll = 0xffffffffu;

__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kAll, v, o));
    return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kAll, v, o);
    return v;
}

__device__ __forceinline__ float ldv(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldv(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// logits: (B*N*fH*fW) channels-last rows of `row_stride` scalars, the first D + 64 are used (fp32, or bf16 straight from the
// tensor-core depthnet whose output rows are padded to 128 channels);  vox: (B*N, D, fH, fW) flat BEV cell or -1
template <bool BACKWARD, typename LT>
__global__ void __launch_bounds__(256)
lift_splat_kernel(const LT* __restrict__ logits, const int* __restrict__ vox, float* __restrict__ bev,
                  const float* __restrict__ g_bev, float* __restrict__ g_logits,
                  int B, int N, int D, int fH, int fW, int XY, int row_stride) {
    constexpr int C = 64;
    const int lane = threadIdx.x & 31;
    const long long n_pix = (long long)B * N * fH * fW;
    const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pix >= n_pix) return;
    const int hw = fH * fW;
    const long long bn = pix / hw;
    const int p_hw = (int)(pix - bn * hw);
    const int b = (int)(bn / N);
    const LT* row = logits + pix * row_stride;

    // depth soft-max (lss.py:60-61,68): lane l owns depth bins l, l+32, ...
    float lg[kMaxDepthSlots], dep[kMaxDepthSlots];
    int vx[kMaxDepthSlots];
    float m = -INFINITY;
#pragma unroll
    for (int s = 0; s < kMaxDepthSlots; ++s) {
        const int d = s * 32 + lane;
        lg[s] = d < D ? ldv(row + d) : -INFINITY;
        vx[s] = d < D ? __ldg(vox + (bn * D + d) * hw + p_hw) : -1;
        m = fmaxf(m, lg[s]);
    }
    m = wmax(m);
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < kMaxDepthSlots; ++s) { dep[s] = s * 32 + lane < D ? __expf(lg[s] - m) : 0.f; sum += dep[s]; }
    sum = wsum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int s = 0; s < kMaxDepthSlots; ++s) dep[s] *= inv;

    // channels 2l, 2l+1 (scalar loads: a row is D + 64 floats and D may be odd, e.g. 59)
    const float2 f = make_float2(ldv(row + D + 2 * lane), ldv(row + D + 2 * lane + 1));

    if (!BACKWARD) {
        float* cellbase = bev + (long long)b * XY * C + 2 * lane;
#pragma unroll
        for (int s = 0; s < kMaxDepthSlots; ++s) {
            const int dmax = min(32, D - s * 32);
            for (int k = 0; k < dmax; ++k) {
                const int v = __shfl_sync(kAll, vx[s], k);
                const float w = __shfl_sync(kAll, dep[s], k);
                if (v >= 0) atomicAdd(reinterpret_cast<float2*>(cellbase + (long long)v * C), make_float2(w * f.x, w * f.y));
            }
        }
    } else {
        const float* gbase = g_bev + (long long)b * XY * C + 2 * lane;
        float gd[kMaxDepthSlots] = {0.f, 0.f, 0.f, 0.f};
        float2 gf = make_float2(0.f, 0.f);
#pragma unroll
        for (int s = 0; s < kMaxDepthSlots; ++s) {
            const int dmax = min(32, D - s * 32);
            for (int k = 0; k < dmax; ++k) {
                const int v = __shfl_sync(kAll, vx[s], k);
                const float w = __shfl_sync(kAll, dep[s], k);
                if (v >= 0) {      // warp-uniform
                    const float2 g = __ldg(reinterpret_cast<const float2*>(gbase + (long long)v * C));
                    gf.x += w * g.x; gf.y += w * g.y;
                    const float dot = wsum(f.x * g.x + f.y * g.y);     // d out / d depth_d
                    if (lane == k) gd[s] = dot;
                }
            }
        }
        // soft-max backward: g_logit_d = depth_d (g_depth_d - sum_d' depth_d' g_depth_d')
        float acc = 0.f;
#pragma unroll
        for (int s = 0; s < kMaxDepthSlots; ++s) acc += dep[s] * gd[s];
        acc = wsum(acc);
        float* grow = g_logits + pix * (D + C);
#pragma unroll
        for (int s = 0; s < kMaxDepthSlots; ++s) {
            const int d = s * 32 + lane;
            if (d < D) grow[d] = dep[s] * (gd[s] - acc);
        }
        grow[D + 2 * lane] = gf.x;
        grow[D + 2 * lane + 1] = gf.y;
    }
}

static const char* check(int B, int N, int D, int C, int fH, int fW, int X, int Y) {
    if (B < 1 || N < 1 || fH < 1 || fW < 1 || X < 1 || Y < 1) return "sizes must be positive";
    if (C != 64) return "the camera feature width must be 64 (LiftSplatShoot.camC, lss.py:182)";
    if (D < 1 || D > 32 * kMaxDepthSlots) return "D must be in [1, 128]";
    return nullptr;
}

}  // namespace mfb

using namespace mfb;

extern "C" {

int mfb_lift_splat_forward(const void* logits, const void* vox, void* bev, int B, int N, int D, int C, int fH, int fW,
                           int X, int Y, void* stream) {
    if (const char* m = check(B, N, D, C, fH, fW, X, Y)) return fail_status(MFB_ERR_INVALID_ARGUMENT, m);
    if (!logits || !vox || !bev) return fail_status(MFB_ERR_INVALID_ARGUMENT, "NULL pointer");
    const long long n_pix = (long long)B * N * fH * fW;
    const int wpb = 8;
    lift_splat_kernel<false, float><<<(unsigned)((n_pix + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        (const float*)logits, (const int*)vox, (float*)bev, nullptr, nullptr, B, N, D, fH, fW, X * Y, D + C);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_status(MFB_ERR_CUDA, std::string("lift_splat forward launch: ") + cudaGetErrorString(e));
    return MFB_OK;
}

int mfb_lift_splat_forward_bf16(const void* logits, int row_stride, const void* vox, void* bev, int B, int N, int D, int C,
                                int fH, int fW, int X, int Y, void* stream) {
    if (const char* m = check(B, N, D, C, fH, fW, X, Y)) return fail_status(MFB_ERR_INVALID_ARGUMENT, m);
    if (!logits || !vox || !bev) return fail_status(MFB_ERR_INVALID_ARGUMENT, "NULL pointer");
    if (row_stride < D + C) return fail_status(MFB_ERR_INVALID_ARGUMENT, "row_stride must be >= D + C");
    const long long n_pix = (long long)B * N * fH * fW;
    const int wpb = 8;
    lift_splat_kernel<false, __nv_bfloat16><<<(unsigned)((n_pix + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)logits, (const int*)vox, (float*)bev, nullptr, nullptr, B, N, D, fH, fW, X * Y, row_stride);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_status(MFB_ERR_CUDA, std::string("lift_splat forward (bf16) launch: ") + cudaGetErrorString(e));
    return MFB_OK;
}

int mfb_lift_splat_backward(const void* logits, const void* vox, const void* g_bev, void* g_logits, int B, int N, int D,
                            int C, int fH, int fW, int X, int Y, void* stream) {
    if (const char* m = check(B, N, D, C, fH, fW, X, Y)) return fail_status(MFB_ERR_INVALID_ARGUMENT, m);
    if (!logits || !vox || !g_bev || !g_logits) return fail_status(MFB_ERR_INVALID_ARGUMENT, "NULL pointer");
    const long long n_pix = (long long)B * N * fH * fW;
    const int wpb = 8;
    lift_splat_kernel<true, float><<<(unsigned)((n_pix + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        (const float*)logits, (const int*)vox, nullptr, (const float*)g_bev, (float*)g_logits, B, N, D, fH, fW, X * Y, D + C);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_status(MFB_ERR_CUDA, std::string("lift_splat backward launch: ") + cudaGetErrorString(e));
    return MFB_OK;
}

}  // extern "C"
