// K7: the memory-bound pieces of the terrain encoder's inference path, NHWC bf16 end to end, so that nothing between the
// tensor-core convolutions (K4, conv_tcgen05.cuh) goes back through NCHW / fp32 / framework ops.
//
//   upsample_concat   Up.forward: cat([skip, bilinear_up(x)], channel)                   terrain_encoder/lss.py:27-46
//                     and the x2 nn.Upsample in front of every BEV head                  lss.py:117-139
//   stem_conv         EfficientNet-B0 stem: 3x3/2 conv (TF "same" padding) + BN + swish  lss.py:78 (efficientnet_pytorch 0.7.1)
//   dwconv            MBConv depthwise k x k conv + BN + swish, with the squeeze-excite
//                     global average pool accumulated on the fly                        lss.py:83-90 (MBConvBlock.forward)
//   se_fold           squeeze-excite MLP (reduce -> swish -> expand -> sigmoid) per image, folded into that image's copy
//                     of the 1x1 projection matrix:  W_n[co,c] = W[co,c] * s_n[c]   (x * s) @ W^T == x @ W_n^T
//   cast              fp32 -> bf16 (the lift-splat BEV grid is accumulated with fp32 atomics)
//
// All of them are HBM-bound: 16-byte vector loads / stores along the channel axis, fp32 arithmetic, one thread per
// (pixel, 8-channel group).  Eval-mode BatchNorm is pre-folded by the host: weights carry the scale, `shift` the rest.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstring>
#include <string>

#include "../../include/monoforce_b200.h"

namespace mfb {
void count_launch();
int fail_status(int code, const std::string& msg);

namespace enc {

struct alignas(16) Bf8 { __nv_bfloat162 v[4]; };

// bf16 -> fp32 is a 16-bit shift: one SHL and one LOP per pair (the library conversion goes through more instructions)
__device__ __forceinline__ void unpack8(const Bf8& b, float* f) {
    const uint4 u = *reinterpret_cast<const uint4*>(&b);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ Bf8 pack8(const float* f) {
    Bf8 b;
#pragma unroll
    for (int i = 0; i < 4; ++i) b.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return b;
}
// acc[0..7] += t[0..7] * w[0..7] as four packed fma.rn.f32x2 (FFMA2: two FMAs per issue slot on sm_100a; the depthwise kernels
// are issue-bound, and the bf16 unpack already leaves channel pairs in adjacent registers)
__device__ __forceinline__ void fma8(float* acc, const float* t, const float* w) {
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const float2 r = __ffma2_rn(make_float2(t[k], t[k + 1]), make_float2(w[k], w[k + 1]), make_float2(acc[k], acc[k + 1]));
        acc[k] = r.x; acc[k + 1] = r.y;
    }
}
__device__ __forceinline__ Bf8 ld8(const __nv_bfloat16* p) { return *reinterpret_cast<const Bf8*>(p); }
// x sigmoid(x) = x/2 (1 + tanh(x/2)): one SFU op + 2 FMA-class instructions (tanh.approx: 2^-11 relative, below bf16 rounding)
__device__ __forceinline__ float silu(float v) {
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// ---------------------------------------------------------------------------------------------------------------
// out[n,h,w, 0:Cs] = skip[n,h,w,:];  out[n,h,w, Cs:Cs+Cl] = bilinear(low)[n,h,w,:] (align_corners=True);  rest = 0
// ---------------------------------------------------------------------------------------------------------------
// IDX = unsigned (the usual case: fewer than 2^32 work items; 64-bit div / mod cost ~10x a 32-bit one) or long long
template <typename IDX>
__global__ void __launch_bounds__(256)
upsample_concat_kernel(const __nv_bfloat16* __restrict__ skip, const __nv_bfloat16* __restrict__ low,
                       __nv_bfloat16* __restrict__ out, int N, int H, int W, int Cs, int Hl, int Wl, int Cl, int Cout,
                       float ry, float rx) {
    const IDX G = (IDX)(Cout >> 3);
    const IDX total = (IDX)N * H * W * G;
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
        const int g = (int)(i % G);
        const IDX pix = i / G;
        const int w = (int)(pix % (IDX)W);
        const IDX nh = pix / (IDX)W;
        const int h = (int)(nh % (IDX)H);
        const int n = (int)(nh / (IDX)H);
        const int c = g << 3;
        Bf8 o;
        if (c < Cs) {
            o = ld8(skip + (long long)pix * Cs + c);
        } else if (c < Cs + Cl) {
            const int cl = c - Cs;
            // torch upsample_bilinear2d, align_corners=True: src = dst * (in - 1) / (out - 1)
            const float sy = ry * h, sx = rx * w;
            const int y0 = min((int)sy, Hl - 1), x0 = min((int)sx, Wl - 1);
            const int y1 = min(y0 + 1, Hl - 1), x1 = min(x0 + 1, Wl - 1);
            const float ly = sy - y0, lx = sx - x0;
            const __nv_bfloat16* base = low + (long long)n * Hl * Wl * Cl + cl;
            float a[8], b[8], cc[8], d[8], r[8];
            unpack8(ld8(base + ((long long)y0 * Wl + x0) * Cl), a);
            unpack8(ld8(base + ((long long)y0 * Wl + x1) * Cl), b);
            unpack8(ld8(base + ((long long)y1 * Wl + x0) * Cl), cc);
            unpack8(ld8(base + ((long long)y1 * Wl + x1) * Cl), d);
            const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = w00 * a[k] + w01 * b[k] + w10 * cc[k] + w11 * d[k];
            o = pack8(r);
        } else {
            const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            o = pack8(z);
        }
        *reinterpret_cast<Bf8*>(out + (long long)pix * Cout + c) = o;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// stem: img (N,3,H,W) fp32 NCHW -> y (N,Ho,Wo,32) bf16 NHWC; 3x3 stride 2, low-side padding (ph, pw); weights (3,3,3,32)
// fp32 [dy][dx][ci][co] with the BN scale folded; y = swish(conv + shift).  The 864 weights + 32 shifts travel BY VALUE as
// a kernel parameter: they sit in the constant bank and feed the FMAs as immediate operands (no load instruction per
// FMA; from shared memory the kernel was LDS-bound: 216 LDS.128 per 864 FMAs, 0.45 ms at 64 x 512 x 512).
// ---------------------------------------------------------------------------------------------------------------
struct StemWeights { float w[27 * 32]; float shift[32]; };

__global__ void __launch_bounds__(128)
stem_conv_kernel(const float* __restrict__ img, const __grid_constant__ StemWeights sw, __nv_bfloat16* __restrict__ y, int N,
                 int H, int W, int Ho, int Wo, int ph, int pw) {
    const unsigned total = (unsigned)N * Ho * Wo;            // host guarantees < 2^31 output pixels
    for (unsigned pix = blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += gridDim.x * blockDim.x) {
        const int wo = (int)(pix % (unsigned)Wo);
        const unsigned nh = pix / (unsigned)Wo;
        const int ho = (int)(nh % (unsigned)Ho);
        const int n = (int)(nh / (unsigned)Ho);
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = sw.shift[c];
        const float* base = img + (long long)n * 3 * H * W;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int hi = 2 * ho + dy - ph;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int wi = 2 * wo + dx - pw;
                const bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float v = ok ? __ldg(base + ((long long)ci * H + hi) * W + wi) : 0.f;
#pragma unroll
                    for (int c = 0; c < 32; ++c) acc[c] = fmaf(v, sw.w[((dy * 3 + dx) * 3 + ci) * 32 + c], acc[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = silu(acc[c]);
        Bf8* dst = reinterpret_cast<Bf8*>(y + (long long)pix * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = pack8(acc + 8 * q);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// depthwise k x k conv + shift + swish; pool[n,c] += sum over the pixels of the output (fp32)
// x (N,H,W,C), y (N,Ho,Wo,C) bf16; w (K*K, C) fp32 (BN scale folded).
// Work item = a strip of kDwStrip consecutive output pixels of one row x one 8-channel group: the input row segment of a
// tap row is loaded once and slides over the strip (k + (S-1)*stride loads per tap row instead of k*S), weights once per
// tap.  A CTA owns all C/8 channel groups of P = blockDim / G strips at a time and walks `rows` strips per thread (up to 8;
// fewer on the small feature maps, so that the grid still fills the GPU a few times over), so its squeeze-excite partial
// sums leave as ONE atomicAdd per channel for up to P * rows * kDwStrip pixels.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDwStrip = 4, kDwMaxRows = 8;

template <int K, int STRIDE>
__global__ void __launch_bounds__(256, 2)
dwconv_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ shift,
              __nv_bfloat16* __restrict__ y, float* __restrict__ pool, int H, int W, int C, int Ho, int Wo, int ph, int pw,
              int rows) {
    extern __shared__ float pool_s[];                      // C partial sums of this CTA
    constexpr int S = kDwStrip;
    constexpr int IN = K + (S - 1) * STRIDE;               // input pixels a strip needs per tap row
    const int n = blockIdx.y;
    const int G = C >> 3;
    const int P = blockDim.x / G;                          // strips in flight per CTA (host guarantees >= 1)
    const int g = threadIdx.x % G, slot = threadIdx.x / G;
    const int c = g << 3;
    const int strips_w = (Wo + S - 1) / S;
    const int n_strips = Ho * strips_w;
    if (pool) {
        for (int i = threadIdx.x; i < C; i += blockDim.x) pool_s[i] = 0.f;
        __syncthreads();
    }
    float sh[8];
    {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(shift + c)), s1 = __ldg(reinterpret_cast<const float4*>(shift + c + 4));
        sh[0] = s0.x; sh[1] = s0.y; sh[2] = s0.z; sh[3] = s0.w; sh[4] = s1.x; sh[5] = s1.y; sh[6] = s1.z; sh[7] = s1.w;
    }
    float psum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const __nv_bfloat16* xin = x + (long long)n * H * W * C + c;
    __nv_bfloat16* yout = y + (long long)n * Ho * Wo * C + c;
    if (slot < P) {
        for (int r = 0; r < rows; ++r) {
            const int sid = (blockIdx.x * rows + r) * P + slot;
            if (sid >= n_strips) break;
            const int ho = sid / strips_w, wo0 = (sid - ho * strips_w) * S;
            float acc[S][8];
#pragma unroll
            for (int j = 0; j < S; ++j)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[j][k] = sh[k];
            const int wi0 = wo0 * STRIDE - pw;
#pragma unroll
            for (int dy = 0; dy < K; ++dy) {
                const int hi = ho * STRIDE + dy - ph;
                if (hi < 0 || hi >= H) continue;
                const __nv_bfloat16* row = xin + (long long)hi * W * C;
                // input-stationary: the row segment is loaded once (packed), every pixel is unpacked ONCE and feeds all the
                // (output j, tap dx) pairs with j*STRIDE + dx == i; the K weights of this tap row sit in registers
                Bf8 v[IN];
                const __nv_bfloat16* p0 = row + (long long)wi0 * C;       // may point before the row: only dereferenced in range
                if (wi0 >= 0 && wi0 + IN <= W) {                          // interior strip: no per-pixel bounds tests
#pragma unroll
                    for (int i = 0; i < IN; ++i) v[i] = ld8(p0 + i * C);
                } else {
#pragma unroll
                    for (int i = 0; i < IN; ++i) {
                        const int wi = wi0 + i;
                        uint4 z = make_uint4(0u, 0u, 0u, 0u);
                        if (wi >= 0 && wi < W) z = *reinterpret_cast<const uint4*>(p0 + i * C);
                        *reinterpret_cast<uint4*>(&v[i]) = z;
                    }
                }
                float wk[K][8];
#pragma unroll
                for (int dx = 0; dx < K; ++dx) {
                    const float* wr = w + (dy * K + dx) * C + c;
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wr)), w1 = __ldg(reinterpret_cast<const float4*>(wr + 4));
                    wk[dx][0] = w0.x; wk[dx][1] = w0.y; wk[dx][2] = w0.z; wk[dx][3] = w0.w;
                    wk[dx][4] = w1.x; wk[dx][5] = w1.y; wk[dx][6] = w1.z; wk[dx][7] = w1.w;
                }
#pragma unroll
                for (int i = 0; i < IN; ++i) {
                    float t[8];
                    unpack8(v[i], t);
#pragma unroll
                    for (int j = 0; j < S; ++j) {
                        const int dx = i - j * STRIDE;
                        if (dx >= 0 && dx < K) {
                            fma8(acc[j], t, wk[dx]);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < S; ++j) {
                if (wo0 + j < Wo) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[j][k] = silu(acc[j][k]);
                    *reinterpret_cast<Bf8*>(yout + ((long long)ho * Wo + wo0 + j) * C) = pack8(acc[j]);
#pragma unroll
                    for (int k = 0; k < 8; ++k) psum[k] += acc[j][k];      // squeeze-excite pool in fp32 (before the bf16 rounding)
                }
            }
        }
    }
    if (pool) {
        if (slot < P) {
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(pool_s + c + k, psum[k]);      // P-way contention at most (P = 256 / G)
        }
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            const float v = pool_s[i];
            if (v != 0.f) atomicAdd(pool + (long long)n * C + i, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Stride-1 depthwise conv through a shared-memory tile (the layers where the strip kernel above was latency-bound: every
// thread waited for its own tap-row loads before its FMAs).  A CTA owns TH x TW output pixels x 32 channels of one image:
// the (TH + K - 1) x (TW + K - 1) input patch is brought in with cp.async (16 bytes per request, zero-filled outside the
// image = the conv's padding), then each thread computes one 4-pixel strip x 8 channels from shared memory.  Several CTAs
// per SM overlap one CTA's loads with another's FMAs.  Pixel pitch in shared memory = 64 + 16 bytes: conflict-free LDS.128.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDwSlab = 32;                      // channels per CTA
constexpr int kDwPitch = kDwSlab * 2 + 16;       // bytes per staged pixel

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;               // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gmem_src), "r"(sz) : "memory");
}

constexpr int kDwChunk = 16;                     // spatial tiles of one (image, channel slab) a CTA works through before it flushes its pool sums

template <int K>
__global__ void __launch_bounds__(256, 2)
dwconv_tile_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ shift,
                   __nv_bfloat16* __restrict__ y, float* __restrict__ pool, int N, int H, int W, int C, int ph, int pw, int TW,
                   int TH) {
    // Persistent and double-buffered: a CTA walks work units = (image, 32-channel slab, chunk of up to 16 spatial tiles); the
    // cp.async requests of tile i+1 are in flight while tile i is computed, and the squeeze-excite partial sums stay in
    // registers for a whole unit (one shuffle-reduce + one atomic per channel per unit instead of per tile).
    extern __shared__ __align__(16) unsigned char dw_smem[];
    constexpr int S = kDwStrip, IN = K + S - 1;
    float* pool_s = reinterpret_cast<float*>(dw_smem);                 // [32]
    const int TWI = TW + K - 1, THI = TH + K - 1;
    const int tile_bytes = THI * TWI * kDwPitch;
    unsigned char* const buf0 = dw_smem + 128;
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    const int sp_tiles = tiles_w * tiles_h;
    const int chunks = (sp_tiles + kDwChunk - 1) / kDwChunk;
    const int slabs = (C + kDwSlab - 1) / kDwSlab;
    const int units = N * slabs * chunks;

    // cp.async staging of spatial tile q of (image n, slab): thread i handles patch entries i, i + 256, ... (entry = pixel * 4 + group)
    const int e_pix0 = threadIdx.x >> 2, e_g = threadIdx.x & 3;
    const int e_iy0 = e_pix0 / TWI, e_ix0 = e_pix0 - e_iy0 * TWI;     // one division per kernel, then incremental
    const int step_iy = 64 / TWI, step_ix = 64 - step_iy * TWI;
    auto stage = [&](int n, int slab, int q, unsigned char* tile) {
        const int h0 = (q / tiles_w) * TH, w0 = (q % tiles_w) * TW, c0 = slab * kDwSlab;
        const bool gok = e_g < min(4, (C - c0) >> 3);
        const __nv_bfloat16* xin = x + (long long)n * H * W * C + c0 + e_g * 8;
        int ix = e_ix0, iy = e_iy0;
        for (int pxy = e_pix0; pxy < THI * TWI; pxy += 64) {
            const int hi = h0 + iy - ph, wi = w0 + ix - pw;
            const bool ok = gok && hi >= 0 && hi < H && wi >= 0 && wi < W;
            cp_async16(tile + pxy * kDwPitch + e_g * 16, ok ? xin + ((long long)hi * W + wi) * C : xin, ok);
            ix += step_ix; iy += step_iy;
            if (ix >= TWI) { ix -= TWI; ++iy; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // thread -> (strip, channel group): 4 groups x (TW/4) strips x TH rows = 256 threads
    const int g = threadIdx.x & 3;
    const int sidx = threadIdx.x >> 2;
    const int sw = sidx % (TW / S), sh = sidx / (TW / S);

    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int n = u / (slabs * chunks), r = u - n * slabs * chunks;
        const int slab = r / chunks, q0 = (r - slab * chunks) * kDwChunk;
        const int q1 = min(q0 + kDwChunk, sp_tiles);
        const int c0 = slab * kDwSlab;
        const int gmax = min(4, (C - c0) >> 3);
        const int c = c0 + g * 8;
        const bool live = g < gmax && sh < TH;
        float sh8[8], psum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (live) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(shift + c)), s1 = __ldg(reinterpret_cast<const float4*>(shift + c + 4));
            sh8[0] = s0.x; sh8[1] = s0.y; sh8[2] = s0.z; sh8[3] = s0.w; sh8[4] = s1.x; sh8[5] = s1.y; sh8[6] = s1.z; sh8[7] = s1.w;
        }
        __syncthreads();                                                    // the previous unit is done with both buffers and pool_s
        if (threadIdx.x < 32) pool_s[threadIdx.x] = 0.f;
        stage(n, slab, q0, buf0);
        for (int q = q0, it = 0; q < q1; ++q, ++it) {
            unsigned char* const tile = buf0 + (it & 1) * tile_bytes;
            if (q + 1 < q1) {
                stage(n, slab, q + 1, buf0 + ((it + 1) & 1) * tile_bytes);      // prefetch the next tile into the other buffer
                asm volatile("cp.async.wait_group 1;" ::: "memory");            // ... and wait for the current one only
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();
            if (live) {
                const int h0 = (q / tiles_w) * TH, w0 = (q % tiles_w) * TW;
                float acc[S][8];
#pragma unroll
                for (int j = 0; j < S; ++j)
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[j][k] = sh8[k];
#pragma unroll
                for (int dy = 0; dy < K; ++dy) {
                    const unsigned char* row = tile + ((sh + dy) * TWI + sw * S) * kDwPitch + g * 16;
                    float wk[K][8];
#pragma unroll
                    for (int dx = 0; dx < K; ++dx) {
                        const float* wr = w + (dy * K + dx) * C + c;
                        const float4 w0v = __ldg(reinterpret_cast<const float4*>(wr)), w1v = __ldg(reinterpret_cast<const float4*>(wr + 4));
                        wk[dx][0] = w0v.x; wk[dx][1] = w0v.y; wk[dx][2] = w0v.z; wk[dx][3] = w0v.w;
                        wk[dx][4] = w1v.x; wk[dx][5] = w1v.y; wk[dx][6] = w1v.z; wk[dx][7] = w1v.w;
                    }
#pragma unroll
                    for (int i = 0; i < IN; ++i) {
                        float tv[8];
                        unpack8(*reinterpret_cast<const Bf8*>(row + i * kDwPitch), tv);
#pragma unroll
                        for (int j = 0; j < S; ++j) {
                            const int dx = i - j;
                            if (dx >= 0 && dx < K) {
fma8(acc[j], tv, wk[dx]);
                            }
                        }
                    }
                }
                const int ho = h0 + sh;
                if (ho < H) {
                    __nv_bfloat16* yout = y + (((long long)n * H + ho) * W) * C + c;
#pragma unroll
                    for (int j = 0; j < S; ++j) {
                        const int wo = w0 + sw * S + j;
                        if (wo < W) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) acc[j][k] = silu(acc[j][k]);
                            *reinterpret_cast<Bf8*>(yout + (long long)wo * C) = pack8(acc[j]);
#pragma unroll
                            for (int k = 0; k < 8; ++k) psum[k] += acc[j][k];
                        }
                    }
                }
            }
            __syncthreads();              // everyone is done with this tile's buffer (it is the prefetch target of the next iteration)
        }
        if (pool) {
            // lanes l, l+4, l+8 ... of a warp hold the same channel group: fold them with xor-shuffles, then 8 shared atomics per group
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float v = psum[k];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                psum[k] = v;
            }
            if ((threadIdx.x & 31) < 4 && g < gmax) {
#pragma unroll
                for (int k = 0; k < 8; ++k) atomicAdd(pool_s + g * 8 + k, psum[k]);
            }
            __syncthreads();
            if (threadIdx.x < 32 && c0 + (int)threadIdx.x < C) {
                const float v = pool_s[threadIdx.x];
                if (v != 0.f) atomicAdd(pool + (long long)n * C + c0 + threadIdx.x, v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// squeeze-excite folded into per-image projection weights, two launches:
//   se_mlp   grid N:  m = pool[n,:] * inv_hw;  r = swish(Wr m + br) (Sq);  s = sigmoid(We r + be) (C);  pool[n,:] <- s
//   se_scale grid (ceil(Cout / kSeRows), N):  out[n,co,c] = proj[co,c] * s[n,c]
// Wr (Sq, Cse), We TRANSPOSED (Sq, Cse): Cse = the block's real channel count; channels [Cse, C) are padding: s = 0 there.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSeRows = 16;
constexpr int kSeMaxC = 1280, kSeMaxSq = 64;
__global__ void __launch_bounds__(256)
se_mlp_kernel(float* __restrict__ pool, float inv_hw, const float* __restrict__ Wr, const float* __restrict__ br,
              const float* __restrict__ We, const float* __restrict__ be, int C, int Cse, int Sq) {
    __shared__ float m[kSeMaxC];
    __shared__ float r[kSeMaxSq];
    float* row = pool + (long long)blockIdx.x * C;
    for (int c = threadIdx.x; c < Cse; c += blockDim.x) m[c] = row[c] * inv_hw;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int q = warp; q < Sq; q += 8) {
        float a = 0.f;
        for (int c = lane; c < Cse; c += 32) a = fmaf(__ldg(Wr + (long long)q * Cse + c), m[c], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) r[q] = silu(a + __ldg(br + q));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float v = 0.f;
        if (c < Cse) {
            float a = __ldg(be + c);
            for (int q = 0; q < Sq; ++q) a = fmaf(__ldg(We + (long long)q * Cse + c), r[q], a);      // We^T (Sq, Cse): coalesced over c
            v = __fdividef(1.f, 1.f + __expf(-a));
        }
        row[c] = v;
    }
}

__global__ void __launch_bounds__(256)
se_scale_kernel(const float* __restrict__ s, const __nv_bfloat16* __restrict__ proj, __nv_bfloat16* __restrict__ out, int C, int Cout) {
    const int n = blockIdx.y;
    const int row0 = blockIdx.x * kSeRows;
    const int G = C >> 3;
    const float* sn = s + (long long)n * C;
    for (int i = threadIdx.x; i < kSeRows * G; i += blockDim.x) {
        const int row = row0 + i / G, c = (i % G) << 3;
        if (row >= Cout) break;
        float v[8];
        unpack8(ld8(proj + (long long)row * C + c), v);
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(sn + c)), s1 = __ldg(reinterpret_cast<const float4*>(sn + c + 4));
        v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w; v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
        *reinterpret_cast<Bf8*>(out + ((long long)n * Cout + row) * C + c) = pack8(v);
    }
}

__global__ void __launch_bounds__(256)
cast_kernel(const float4* __restrict__ src, Bf8* __restrict__ dst, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(src + 2 * i), b = __ldg(src + 2 * i + 1);
        const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        dst[i] = pack8(f);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Between the encoder and the physics: terrain = geom - diff (lss.py:158) and the AvgPool2d(k) that brings the
// encoder's grid to the physics resolution (train.py:96-99,234-235), for the two maps the rollout reads, in one pass.
// geom / diff / friction: (B,1,X,Y) fp32 with batch stride `bs` (they may be channel slices of one (B,3,X,Y) tensor).
// One thread per pooled cell: writes its k x k terrain cells and the two pooled means.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
terrain_postproc_kernel(const float* __restrict__ geom, const float* __restrict__ diff, const float* __restrict__ fric,
                        long long bs, float* __restrict__ terrain, float* __restrict__ z_pool, float* __restrict__ mu_pool,
                        int B, int X, int Y, int k) {
    const int Xp = X / k, Yp = Y / k;
    const long long total = (long long)B * Xp * Yp;
    const float inv = 1.f / (float)(k * k);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int yp = (int)(i % Yp);
        const int xp = (int)((i / Yp) % Xp);
        const int b = (int)(i / ((long long)Yp * Xp));
        float sz = 0.f, sm = 0.f;
        for (int dx = 0; dx < k; ++dx) {
            const long long row = (long long)(xp * k + dx) * Y + yp * k;
            for (int dy = 0; dy < k; ++dy) {
                const float t = __ldg(geom + b * bs + row + dy) - __ldg(diff + b * bs + row + dy);
                if (terrain) terrain[(long long)b * X * Y + row + dy] = t;
                sz += t;
                sm += __ldg(fric + b * bs + row + dy);
            }
        }
        if (z_pool) z_pool[i] = sz * inv;
        if (mu_pool) mu_pool[i] = sm * inv;
    }
    // rows / columns beyond the last full window are dropped by AvgPool2d but still belong to `terrain`
    if (terrain && (X % k || Y % k)) {
        const long long cells = (long long)B * X * Y;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x) {
            const int y = (int)(i % Y), x = (int)((i / Y) % X);
            if (x >= Xp * k || y >= Yp * k) {
                const int b = (int)(i / ((long long)X * Y));
                const long long o = (long long)x * Y + y;
                terrain[i] = __ldg(geom + b * bs + o) - __ldg(diff + b * bs + o);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Planner post-processing of a rollout: 4x4 poses (monoforce_node.py:80-85, diff_physics.py:246-250) and the
// inclination cost mean_t |roll| + mean_t |pitch| with (roll, pitch) the extrinsic x-y-z Euler angles of R
// (diff_physics.py:262-266: scipy Rotation.from_matrix(R).as_euler('xyz')): roll = atan2(R21, R22), pitch = -asin(R20).
// One warp per trajectory.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
path_postproc_kernel(const float* __restrict__ Xs, const float* __restrict__ Rs, float* __restrict__ poses,
                     float* __restrict__ cost, int B, int T) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    float acc = 0.f;
    for (int t = lane; t < T; t += 32) {
        const float* R = Rs + ((long long)b * T + t) * 9;
        const float* x = Xs + ((long long)b * T + t) * 3;
        float r[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) r[i] = __ldg(R + i);
        if (poses) {
            float4* P = reinterpret_cast<float4*>(poses + ((long long)b * T + t) * 16);
            P[0] = make_float4(r[0], r[1], r[2], __ldg(x));
            P[1] = make_float4(r[3], r[4], r[5], __ldg(x + 1));
            P[2] = make_float4(r[6], r[7], r[8], __ldg(x + 2));
            P[3] = make_float4(0.f, 0.f, 0.f, 1.f);
        }
        acc += fabsf(atan2f(r[7], r[8])) + fabsf(asinf(fminf(fmaxf(-r[6], -1.f), 1.f)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (cost && lane == 0) cost[b] = acc / (float)T;
}

static int after_launch(const char* what) {
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_status(MFB_ERR_CUDA, std::string(what) + " launch: " + cudaGetErrorString(e));
    return MFB_OK;
}
static unsigned grid_for(long long work, int block, int max_ctas = 148 * 16) {
    long long g = (work + block - 1) / block;
    return (unsigned)(g < 1 ? 1 : (g > max_ctas ? max_ctas : g));
}

}  // namespace enc
}  // namespace mfb

using namespace mfb;
using namespace mfb::enc;

extern "C" {

int mfb_upsample_concat_nhwc_bf16(const void* skip, const void* low, void* out, int N, int H, int W, int C_skip, int Hl,
                                  int Wl, int C_low, int C_out, void* stream) {
    if (!low || !out || (C_skip > 0 && !skip)) return fail_status(MFB_ERR_INVALID_ARGUMENT, "upsample_concat: NULL pointer");
    if (N < 1 || H < 1 || W < 1 || Hl < 1 || Wl < 1) return fail_status(MFB_ERR_INVALID_ARGUMENT, "upsample_concat: sizes must be positive");
    if ((C_skip | C_low | C_out) & 7 || C_low < 8 || C_skip < 0 || C_out < C_skip + C_low)
        return fail_status(MFB_ERR_UNSUPPORTED, "upsample_concat: channel counts must be multiples of 8 and C_out >= C_skip + C_low");
    if (((uintptr_t)skip | (uintptr_t)low | (uintptr_t)out) & 15) return fail_status(MFB_ERR_INVALID_ARGUMENT, "upsample_concat: tensors must be 16-byte aligned");
    const float ry = H > 1 ? (float)(Hl - 1) / (float)(H - 1) : 0.f, rx = W > 1 ? (float)(Wl - 1) / (float)(W - 1) : 0.f;
    const long long total = (long long)N * H * W * (C_out >> 3);
    if (total < (1ll << 31))
        upsample_concat_kernel<unsigned><<<grid_for(total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)skip, (const __nv_bfloat16*)low, (__nv_bfloat16*)out, N, H, W, C_skip, Hl, Wl, C_low, C_out, ry, rx);
    else
        upsample_concat_kernel<long long><<<grid_for(total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)skip, (const __nv_bfloat16*)low, (__nv_bfloat16*)out, N, H, W, C_skip, Hl, Wl, C_low, C_out, ry, rx);
    return after_launch("upsample_concat");
}

int mfb_stem_conv_bf16(const void* img, const void* w, const void* shift, void* y, int N, int H, int W, int Ho, int Wo,
                       int pad_h, int pad_w, void* stream) {
    if (!img || !w || !shift || !y) return fail_status(MFB_ERR_INVALID_ARGUMENT, "stem_conv: NULL pointer");
    if (N < 1 || H < 1 || W < 1 || Ho < 1 || Wo < 1) return fail_status(MFB_ERR_INVALID_ARGUMENT, "stem_conv: sizes must be positive");
    if (2 * (Ho - 1) - pad_h >= H || 2 * (Wo - 1) - pad_w >= W) return fail_status(MFB_ERR_INVALID_ARGUMENT, "stem_conv: output size does not fit");
    if ((long long)N * Ho * Wo >= (1ll << 31)) return fail_status(MFB_ERR_UNSUPPORTED, "stem_conv: more than 2^31 output pixels");
    const long long total = (long long)N * Ho * Wo;
    StemWeights sw;                                          // HOST pointers: the weights are passed by value
    memcpy(sw.w, w, sizeof(sw.w));
    memcpy(sw.shift, shift, sizeof(sw.shift));
    stem_conv_kernel<<<grid_for(total, 128, 148 * 16), 128, 0, (cudaStream_t)stream>>>((const float*)img, sw, (__nv_bfloat16*)y, N, H, W,
                                                                                       Ho, Wo, pad_h, pad_w);
    return after_launch("stem_conv");
}

int mfb_dwconv_bn_silu_bf16(const void* x, const void* w, const void* shift, void* y, void* pool, int N, int H, int W, int C,
                            int Ho, int Wo, int K, int stride, int pad_h, int pad_w, void* stream) {
    if (!x || !w || !shift || !y) return fail_status(MFB_ERR_INVALID_ARGUMENT, "dwconv: NULL pointer");
    if (N < 1 || H < 1 || W < 1 || Ho < 1 || Wo < 1) return fail_status(MFB_ERR_INVALID_ARGUMENT, "dwconv: sizes must be positive");
    if (C & 7 || C < 8 || C > 12288) return fail_status(MFB_ERR_UNSUPPORTED, "dwconv: C must be a multiple of 8");
    if (K != 3 && K != 5) return fail_status(MFB_ERR_UNSUPPORTED, "dwconv: kernel size must be 3 or 5");
    if (stride != 1 && stride != 2) return fail_status(MFB_ERR_UNSUPPORTED, "dwconv: stride must be 1 or 2");
    if (((uintptr_t)x | (uintptr_t)y | (uintptr_t)w | (uintptr_t)shift) & 15) return fail_status(MFB_ERR_INVALID_ARGUMENT, "dwconv: tensors must be 16-byte aligned");
    if (stride == 1 && Ho == H && Wo == W) {
        // stride 1 ("same"): shared-memory tile kernel; 32-wide tiles unless the map is narrower
        const int TW = W > 16 ? 32 : 16, TH = 256 / TW;                 // 4 groups x (TW/4) strips x TH rows = 256 threads
        const long long sp = (long long)((W + TW - 1) / TW) * ((H + TH - 1) / TH);
        const long long tiles = ((sp + kDwChunk - 1) / kDwChunk) * ((C + kDwSlab - 1) / kDwSlab) * N;     // work units
        if (tiles >= (1ll << 31)) return fail_status(MFB_ERR_UNSUPPORTED, "dwconv: too many tiles");
        const long long resident = 148ll * 2;                          // persistent: 2 CTAs per SM
        dim3 grid((unsigned)(tiles < resident ? tiles : resident));
        const size_t smem = 128 + 2 * (size_t)(TH + K - 1) * (TW + K - 1) * kDwPitch;    // two tile buffers
        auto go = [&](auto kern) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
            kern<<<grid, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const float*)w, (const float*)shift,
                                                           (__nv_bfloat16*)y, (float*)pool, N, H, W, C, pad_h, pad_w, TW, TH);
            return true;
        };
        if (!(K == 3 ? go(dwconv_tile_kernel<3>) : go(dwconv_tile_kernel<5>)))
            return fail_status(MFB_ERR_CUDA, "dwconv: cudaFuncSetAttribute failed");
        return after_launch("dwconv (tile)");
    }
    const int G = C >> 3;
    if (G > 256) return fail_status(MFB_ERR_UNSUPPORTED, "dwconv: at most 2048 channels");
    const int P = 256 / G;
    const int threads = ((P * G + 31) / 32) * 32;
    const long long n_strips = (long long)Ho * ((Wo + kDwStrip - 1) / kDwStrip);
    // strips per thread: as many as possible (fewer pool atomics) while the grid keeps >= 4 waves of 2 CTAs per SM
    long long rows = (n_strips * N) / ((long long)P * 148 * 2 * 4);
    rows = rows < 1 ? 1 : (rows > kDwMaxRows ? kDwMaxRows : rows);
    dim3 grid((unsigned)((n_strips + (long long)P * rows - 1) / ((long long)P * rows)), (unsigned)N);
    const size_t smem = (size_t)C * sizeof(float);
    auto go = [&](auto kern) {
        kern<<<grid, threads, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const float*)w, (const float*)shift,
                                                            (__nv_bfloat16*)y, (float*)pool, H, W, C, Ho, Wo, pad_h, pad_w, (int)rows);
    };
    if (K == 3 && stride == 1) go(dwconv_kernel<3, 1>);
    else if (K == 3) go(dwconv_kernel<3, 2>);
    else if (stride == 1) go(dwconv_kernel<5, 1>);
    else go(dwconv_kernel<5, 2>);
    return after_launch("dwconv");
}

int mfb_se_fold_bf16(void* pool, float inv_hw, const void* w_reduce, const void* b_reduce, const void* w_expand,
                     const void* b_expand, const void* proj_w, void* out_w, int N, int C, int C_se, int Sq, int Cout, void* stream) {
    if (!pool || !w_reduce || !b_reduce || !w_expand || !b_expand || !proj_w || !out_w) return fail_status(MFB_ERR_INVALID_ARGUMENT, "se_fold: NULL pointer");
    if (N < 1 || Cout < 1 || Sq < 1 || Sq > kSeMaxSq || C < 8 || C > kSeMaxC || (C & 7) || C_se < 1 || C_se > C)
        return fail_status(MFB_ERR_UNSUPPORTED, "se_fold: need C % 8 == 0, C <= 1280, Sq <= 64, C_se <= C");
    if ((uintptr_t)pool & 15) return fail_status(MFB_ERR_INVALID_ARGUMENT, "se_fold: pool must be 16-byte aligned");
    se_mlp_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>((float*)pool, inv_hw, (const float*)w_reduce, (const float*)b_reduce,
                                                                 (const float*)w_expand, (const float*)b_expand, C, C_se, Sq);
    if (int rc = after_launch("se_mlp")) return rc;
    dim3 grid((unsigned)((Cout + kSeRows - 1) / kSeRows), (unsigned)N);
    se_scale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)pool, (const __nv_bfloat16*)proj_w, (__nv_bfloat16*)out_w, C, Cout);
    return after_launch("se_scale");
}

int mfb_terrain_postproc(const void* geom, const void* diff, const void* friction, long long batch_stride, void* terrain,
                         void* z_pooled, void* mu_pooled, int B, int X, int Y, int k, void* stream) {
    if (!geom || !diff || !friction) return fail_status(MFB_ERR_INVALID_ARGUMENT, "terrain_postproc: NULL pointer");
    if (B < 1 || X < 1 || Y < 1 || k < 1 || k > X || k > Y || batch_stride < (long long)X * Y)
        return fail_status(MFB_ERR_INVALID_ARGUMENT, "terrain_postproc: bad sizes");
    const long long total = (long long)B * (X / k) * (Y / k);
    terrain_postproc_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float*)geom, (const float*)diff, (const float*)friction, batch_stride, (float*)terrain, (float*)z_pooled,
        (float*)mu_pooled, B, X, Y, k);
    return after_launch("terrain_postproc");
}

int mfb_path_postproc(const void* Xs, const void* Rs, void* poses, void* cost, int B, int T, void* stream) {
    if (!Xs || !Rs || (!poses && !cost)) return fail_status(MFB_ERR_INVALID_ARGUMENT, "path_postproc: NULL pointer");
    if (B < 1 || T < 1) return fail_status(MFB_ERR_INVALID_ARGUMENT, "path_postproc: sizes must be positive");
    if ((uintptr_t)poses & 15) return fail_status(MFB_ERR_INVALID_ARGUMENT, "path_postproc: poses must be 16-byte aligned");
    path_postproc_kernel<<<(unsigned)((B + 3) / 4), 128, 0, (cudaStream_t)stream>>>((const float*)Xs, (const float*)Rs, (float*)poses,
                                                                                    (float*)cost, B, T);
    return after_launch("path_postproc");
}

int mfb_cast_f32_to_bf16(const void* src, void* dst, long long n, void* stream) {
    if (!src || !dst) return fail_status(MFB_ERR_INVALID_ARGUMENT, "cast: NULL pointer");
    if (n < 8 || (n & 7) || (((uintptr_t)src | (uintptr_t)dst) & 15)) return fail_status(MFB_ERR_UNSUPPORTED, "cast: n must be a multiple of 8, pointers 16-byte aligned");
    cast_kernel<<<grid_for(n >> 3, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>((const float4*)src, (Bf8*)dst, n >> 3);
    return after_launch("cast");
}

}  // extern "C"
