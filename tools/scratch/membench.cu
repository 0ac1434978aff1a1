// Scratch micro-benchmark (not product): how fast can B200 absorb the rollout's output pattern?
// pattern A: one warp per trajectory writes T rows of ROW floats to two (B,T,ROW) tensors, all warps in lock step
// pattern B: the same bytes written as one linear stream
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(128, 4) rows_stg(float* A, float* Bf, int B, int T, int ROW, int spin) {
    const int lane = threadIdx.x & 31, b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    float acc = lane;
    for (int t = 0; t < T; ++t) {
        for (int k = 0; k < spin; ++k) acc = acc * 1.0001f + 0.5f;     // stand-in for compute
        float* a = A + ((long long)b * T + t) * ROW;
        float* f = Bf + ((long long)b * T + t) * ROW;
        for (int i = lane; i < ROW; i += 32) { a[i] = acc; f[i] = acc; }
    }
}

__global__ void __launch_bounds__(128, 4) rows_tma(float* A, float* Bf, int B, int T, int ROW, int spin, int chunk) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, b = blockIdx.x * 4 + w;
    if (b >= B) return;
    float* img = sm + w * (chunk * ROW + 8);
    float acc = lane;
    // rows are written `chunk` steps at a time (chunk*ROW*4 bytes must be a multiple of 16 and 16B-aligned start)
    for (int t = 0; t < T; t += chunk) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        for (int k = 0; k < spin * chunk; ++k) acc = acc * 1.0001f + 0.5f;
        for (int i = lane; i < ROW * chunk; i += 32) img[i] = acc;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            uint32_t s = (uint32_t)__cvta_generic_to_shared(img);
            uint32_t bytes = chunk * ROW * 4;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(A + ((long long)b * T + t) * ROW), "r"(s), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(Bf + ((long long)b * T + t) * ROW), "r"(s), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void linear(float4* p, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

int main() {
    const int B = 4096, T = 400, ROW = 672;     // 672 = 669 rounded to a 16-byte multiple (keeps TMA alignment trivial)
    const long long n = (long long)B * T * ROW;
    float *A, *Bf;
    CK(cudaMalloc(&A, n * 4)); CK(cudaMalloc(&Bf, n * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto fn) {
        float best = 1e9;
        for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); fn(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
        printf("%-46s %8.3f ms  %7.0f GB/s\n", name, best, 2.0 * n * 4 / best / 1e6);
        cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(e));
    };
    timeit("linear float4 stream (2 x 4.4 GB)", [&] { linear<<<148 * 8, 256>>>((float4*)A, n / 4); linear<<<148 * 8, 256>>>((float4*)Bf, n / 4); });
    for (int spin : {0, 200, 600, 1200}) {
        char nm[96]; snprintf(nm, 96, "rows STG.32 coalesced, spin=%d", spin);
        timeit(nm, [&] { rows_stg<<<B / 4, 128>>>(A, Bf, B, T, ROW, spin); });
    }
    for (int chunk : {1, 2, 4}) for (int spin : {0, 600}) {
        char nm[96]; snprintf(nm, 96, "rows TMA bulk, chunk=%d steps, spin=%d", chunk, spin);
        size_t smem = 4 * (chunk * ROW + 8) * 4;
        cudaFuncSetAttribute(rows_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        timeit(nm, [&] { rows_tma<<<B / 4, 128, smem>>>(A, Bf, B, T, ROW, spin, chunk); });
    }
    return 0;
}
