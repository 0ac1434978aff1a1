// Scratch micro-benchmark: does fma.rn.f32x2 (FFMA2) free issue slots on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, int* iout, float a, float b, int m) {
    float x[16]; int q[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = threadIdx.x + i;
    for (int it = 0; it < N_IT; ++it) {
        if (MODE == 0 || MODE == 2) {          // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        } else {                                // 8 packed FFMA2
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float2 v = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a), make_float2(b, b));
                x[i] = v.x; x[i + 1] = v.y;
            }
        }
        if (MODE >= 2) {                        // plus 16 ALU-pipe integer ops
#pragma unroll
            for (int i = 0; i < 8; ++i) { q[i] = (q[i] ^ m) + it; }
        }
    }
    float s = 0; int t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) t += q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
int main() {
    float* o; int* io; cudaMalloc(&o, 148 * 16 * 128 * 4); cudaMalloc(&io, 148 * 16 * 128 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[4] = {"16 FFMA", "8 FFMA2", "16 FFMA + 8x(xor,add)", "8 FFMA2 + 8x(xor,add)"};
    for (int blocks_per_sm : {1, 2, 4, 8}) for (int mode = 0; mode < 4; ++mode) {
        float best = 1e9;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * blocks_per_sm, 128>>>(o, io, 1.0001f, 0.5f, 12345);
            if (mode == 1) k<1><<<148 * blocks_per_sm, 128>>>(o, io, 1.0001f, 0.5f, 12345);
            if (mode == 2) k<2><<<148 * blocks_per_sm, 128>>>(o, io, 1.0001f, 0.5f, 12345);
            if (mode == 3) k<3><<<148 * blocks_per_sm, 128>>>(o, io, 1.0001f, 0.5f, 12345);
            cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double cyc = best * 1e-3 * 1.965e9 / N_IT;
        printf("warps/SMSP=%d  %-20s %7.3f ms  ~%6.1f cycles/iter  (%.2f cyc per warp-iter per SMSP)\n", blocks_per_sm, names[mode], best, cyc, cyc / blocks_per_sm);
    }
    return 0;
}
