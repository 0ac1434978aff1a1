// K7: the memory-bound pieces of the terrain encoder's inference path, NHWC bf16 end to end, so that nothing between the
// tensor-core convolutions (K4, conv_tcgen05.cuh) goes back through NCHW / fp32 / framework ops.
//
//   upsample_concat   Up.forward: cat([skip, bilinear_up(x)], channel)                   terrain_encoder/lss.py:27-46
//                     and the x2 nn.Upsample in front of every BEV head                  lss.py:117-139
//   stem_conv         EfficientNet-B0 stem: 3x3/2 conv (TF "same" padding) + BN + swish  lss.py:78 (efficientnet_pytorch 0.7.1)
//   dwconv            MBConv depthwise k x k conv + BN + swish, with the squeeze-excite
//                     global average pool accumulated on the fly (kernel: dwconv_tma.cu)  lss.py:83-90 (MBConvBlock.forward)
//   se_fold           squeeze-excite MLP (reduce -> swish -> expand -> sigmoid) per image, folded into that image's copy
//                     of the 1x1 projection matrix:  W_n[co,c] = W[co,c] * s_n[c]   (x * s) @ W^T == x @ W_n^T
//   cast              fp32 -> bf16 (the lift-splat BEV grid is accumulated with fp32 atomics)
//
// Apart from the depthwise convolution (fp32-FMA-bound, its own file) they are HBM-bound: 16-byte vector loads / stores along
// the channel axis, fp32 arithmetic, one thread per (pixel, 8-channel group).  Eval-mode BatchNorm is pre-folded by the host: weights carry the scale, `shift` the rest.
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
int launch_dwconv_tma(const void* x, const float* w, const float* shift, void* y, float* pool, int N, int H, int W, int C, int Ho,
                      int Wo, int K, int stride, int pad_h, int pad_w, cudaStream_t st);
}

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
// Work item = one SOURCE cell (n, y0, x0) x one 8-channel group: the cell's four corner pixels are loaded and unpacked once
// and every output pixel whose interpolation footprint is that cell (4 on average for x2, 16 for x4) is produced from
// registers.  The output-pixel version (4 loads + 32 unpack instructions per 16 output bytes) ran the x2 up-sample in front of
// the BEV heads (0.54 GB out) at 1.8 TB/s with the load/store queue as the top stall (ncu: lg_throttle 8.6 cycles / issue,
// 217 instructions per 16 bytes).  Cell membership uses the forward mapping itself, src = min(int(r * dst), in - 1) (torch
// upsample_bilinear2d, align_corners=True), so every output pixel is written exactly once.
__device__ __forceinline__ int src_index(int dst, float r, int n_in) { return min((int)(r * dst), n_in - 1); }

// smallest dst in [0, n_out] with src_index(dst) >= s (n_out if none): src_index is monotone in dst
__device__ __forceinline__ int first_dst(int s, float r, int n_out, int n_in) {
    if (s <= 0) return 0;
    if (r <= 0.f) return n_out;
    int d = min(max((int)ceilf((float)s / r), 0), n_out);
    while (d > 0 && src_index(d - 1, r, n_in) >= s) --d;
    while (d < n_out && src_index(d, r, n_in) < s) ++d;
    return d;
}

template <typename IDX>
__global__ void __launch_bounds__(256)
upsample_concat_kernel(const __nv_bfloat16* __restrict__ skip, const __nv_bfloat16* __restrict__ low,
                       __nv_bfloat16* __restrict__ out, int N, int H, int W, int Cs, int Hl, int Wl, int Cl, int Cout,
                       float ry, float rx) {
    const IDX G = (IDX)(Cout >> 3);
    const IDX total = (IDX)N * Hl * Wl * G;
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
        const int g = (int)(i % G);
        const IDX cell = i / G;
        const int x0 = (int)(cell % (IDX)Wl);
        const IDX ny = cell / (IDX)Wl;
        const int y0 = (int)(ny % (IDX)Hl);
        const int n = (int)(ny / (IDX)Hl);
        const int c = g << 3;
        // output pixels owned by this cell
        const int h_lo = first_dst(y0, ry, H, Hl), h_hi = y0 + 1 < Hl ? first_dst(y0 + 1, ry, H, Hl) : H;
        const int w_lo = first_dst(x0, rx, W, Wl), w_hi = x0 + 1 < Wl ? first_dst(x0 + 1, rx, W, Wl) : W;
        if (h_lo >= h_hi || w_lo >= w_hi) continue;
        __nv_bfloat16* const obase = out + (long long)n * H * W * Cout + c;
        if (c < Cs) {
            const __nv_bfloat16* sbase = skip + (long long)n * H * W * Cs + c;
            for (int h = h_lo; h < h_hi; ++h)
                for (int w = w_lo; w < w_hi; ++w)
                    *reinterpret_cast<Bf8*>(obase + ((long long)h * W + w) * Cout) = ld8(sbase + ((long long)h * W + w) * Cs);
        } else if (c < Cs + Cl) {
            const int y1 = min(y0 + 1, Hl - 1), x1 = min(x0 + 1, Wl - 1);
            const __nv_bfloat16* base = low + (long long)n * Hl * Wl * Cl + (c - Cs);
            float a[8], b[8], cc[8], d[8];
            unpack8(ld8(base + ((long long)y0 * Wl + x0) * Cl), a);
            unpack8(ld8(base + ((long long)y0 * Wl + x1) * Cl), b);
            unpack8(ld8(base + ((long long)y1 * Wl + x0) * Cl), cc);
            unpack8(ld8(base + ((long long)y1 * Wl + x1) * Cl), d);
            for (int h = h_lo; h < h_hi; ++h) {
                const float ly = ry * h - y0;
                __nv_bfloat16* orow = obase + (long long)h * W * Cout;
                for (int w = w_lo; w < w_hi; ++w) {
                    const float lx = rx * w - x0;
                    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
                    float r[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) r[k] = w00 * a[k] + w01 * b[k] + w10 * cc[k] + w11 * d[k];
                    *reinterpret_cast<Bf8*>(orow + (long long)w * Cout) = pack8(r);
                }
            }
        } else {
            const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const Bf8 zero = pack8(z);
            for (int h = h_lo; h < h_hi; ++h)
                for (int w = w_lo; w < w_hi; ++w) *reinterpret_cast<Bf8*>(obase + ((long long)h * W + w) * Cout) = zero;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// stem: img (N,3,H,W) fp32 NCHW -> y (N,Ho,Wo,32) bf16 NHWC; 3x3 stride 2, low-side padding (ph, pw); weights (3,3,3,32)
// fp32 [dy][dx][ci][co] with the BN scale folded; y = swish(conv + shift).
// As a GEMM the layer is (pixels) x 27 x 32.  On the fp32 pipe it needs 864 FMAs per pixel and was FMA-bound (0.33 ms at
// 64 x 512 x 512, where its 0.47 GB of traffic take 0.08 ms), so the contraction runs on the tensor cores: a warp gathers the
// im2col rows of 16 consecutive output pixels straight into mma.sync.m16n8k16 bf16 A fragments (K padded 27 -> 32, image and
// weights rounded to bf16 like every other layer of the trunk, fp32 accumulate), 8 MMAs per 16 pixels.  mma.sync rather than
// tcgen05 on purpose: K = 27 and N = 32 make the tensor op ~2 % of the kernel; the work is the gather and the 64-byte rows.
// The weights travel BY VALUE as a kernel parameter (2 KB bf16 + shifts): every lane reads its B fragments from the constant
// bank once.  Output rows of the 16 pixels are contiguous (1 KB): they leave through a per-warp staging tile as 16-byte stores.
// ---------------------------------------------------------------------------------------------------------------
struct StemWeights { uint16_t w[32][32]; float shift[32]; };          // w[k][co] bf16 bits, k = (dy * 3 + dx) * 3 + ci, rows 27..31 zero
constexpr int kStemWarps = 8;

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

__global__ void __launch_bounds__(kStemWarps * 32)
stem_conv_kernel(const float* __restrict__ img, const __grid_constant__ StemWeights sw, __nv_bfloat16* __restrict__ y, int N,
                 int H, int W, int Ho, int Wo, int ph, int pw) {
    __shared__ __align__(16) uint32_t stage[kStemWarps][16][17];       // 16 pixels x 32 bf16 (16 words), +1 word: conflict-free
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const unsigned total = (unsigned)N * Ho * Wo;                       // host guarantees < 2^31 output pixels
    const unsigned tiles = (total + 15) / 16;
    const long long HW = (long long)H * W;

    // this lane's 8 im2col columns: k = 2t, 2t+1, 2t+8, 2t+9 (+16 for the second k-step); offset / tap position of each
    int koff[8], kdy[8], kdx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = 2 * t + (j & 1) + ((j >> 1) & 1) * 8 + (j >> 2) * 16;
        const int tap = k / 3, ci = k - tap * 3, dy = tap / 3, dx = tap - dy * 3;
        const bool real = k < 27;
        kdy[j] = real ? dy : 0; kdx[j] = real ? dx : 0;                  // padded columns read a valid pixel; their weights are 0
        koff[j] = real ? (int)(ci * HW) + dy * W + dx : 0;
    }
    // B fragments: b[ks][nt] = {B[16 ks + 2t][8 nt + g], B[.. + 1][..]}, {B[16 ks + 2t + 8][..], B[.. + 9][..]}
    uint32_t bfrag[2][4][2];
    float shv[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int k0 = 16 * ks + 2 * t, n = 8 * nt + g;
            bfrag[ks][nt][0] = (uint32_t)sw.w[k0][n] | ((uint32_t)sw.w[k0 + 1][n] << 16);
            bfrag[ks][nt][1] = (uint32_t)sw.w[k0 + 8][n] | ((uint32_t)sw.w[k0 + 9][n] << 16);
        }
        shv[nt][0] = sw.shift[8 * nt + 2 * t];
        shv[nt][1] = sw.shift[8 * nt + 2 * t + 1];
    }

    for (unsigned tile = blockIdx.x * kStemWarps + warp; tile < tiles; tile += gridDim.x * kStemWarps) {
        // rows g and g + 8 of the tile = output pixels p0, p1
        float av[2][8];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            unsigned pix = tile * 16 + g + 8 * r;
            const bool live = pix < total;
            pix = live ? pix : total - 1;
            const int wo = (int)(pix % (unsigned)Wo);
            const unsigned nh = pix / (unsigned)Wo;
            const int ho = (int)(nh % (unsigned)Ho), n = (int)(nh / (unsigned)Ho);
            const int iy0 = 2 * ho - ph, ix0 = 2 * wo - pw;
            const float* base = img + (long long)n * 3 * HW + (long long)iy0 * W + ix0;
            if (iy0 >= 0 && iy0 + 2 < H && ix0 >= 0 && ix0 + 2 < W) {
#pragma unroll
                for (int j = 0; j < 8; ++j) av[r][j] = __ldg(base + koff[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool ok = (unsigned)(iy0 + kdy[j]) < (unsigned)H && (unsigned)(ix0 + kdx[j]) < (unsigned)W;
                    av[r][j] = ok ? __ldg(base + koff[j]) : 0.f;
                }
            }
        }
        float d[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { d[nt][0] = d[nt][2] = shv[nt][0]; d[nt][1] = d[nt][3] = shv[nt][1]; }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const uint32_t a0 = pack_bf16x2(av[0][4 * ks + 0], av[0][4 * ks + 1]), a1 = pack_bf16x2(av[1][4 * ks + 0], av[1][4 * ks + 1]);
            const uint32_t a2 = pack_bf16x2(av[0][4 * ks + 2], av[0][4 * ks + 3]), a3 = pack_bf16x2(av[1][4 * ks + 2], av[1][4 * ks + 3]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[nt][0]), "+f"(d[nt][1]), "+f"(d[nt][2]), "+f"(d[nt][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bfrag[ks][nt][0]), "r"(bfrag[ks][nt][1]));
        }
        // swish, bf16, through the staging tile: row = pixel of the tile, word = channel pair
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            stage[warp][g][4 * nt + t] = pack_bf16x2(silu(d[nt][0]), silu(d[nt][1]));
            stage[warp][g + 8][4 * nt + t] = pack_bf16x2(silu(d[nt][2]), silu(d[nt][3]));
        }
        __syncwarp();
        // 16 pixels x 64 bytes are contiguous in y: 64 16-byte pieces, two per lane
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int piece = lane + 32 * h, row = piece >> 2, q = piece & 3;
            const unsigned pix = tile * 16 + row;
            if (pix < total) {
                uint4 v;
                v.x = stage[warp][row][4 * q + 0]; v.y = stage[warp][row][4 * q + 1];
                v.z = stage[warp][row][4 * q + 2]; v.w = stage[warp][row][4 * q + 3];
                *reinterpret_cast<uint4*>(y + (long long)pix * 32 + q * 8) = v;
            }
        }
        __syncwarp();
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
// The two matrix-vector products are latency-bound (one CTA per image, ~0.4 MB of weights from L2 / HBM, a few images in
// flight): what counts is the number of loads in flight, so the CTA is 1024 threads and every lane issues 16-byte loads
// back to back (the first version, 256 threads and one scalar load per fused multiply-add, took 46-66 us per call at
// C = 1152).
constexpr int kSeThreads = 1024;
__global__ void __launch_bounds__(kSeThreads)
se_mlp_kernel(float* __restrict__ pool, float inv_hw, const float* __restrict__ Wr, const float* __restrict__ br,
              const float* __restrict__ We, const float* __restrict__ be, int C, int Cse, int Sq) {
    __shared__ __align__(16) float m[kSeMaxC];
    __shared__ float r[kSeMaxSq];
    float* row = pool + (long long)blockIdx.x * C;
    for (int c = threadIdx.x; c < Cse; c += kSeThreads) m[c] = row[c] * inv_hw;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // squeeze: one row of Wr per warp
    for (int q = warp; q < Sq; q += kSeThreads / 32) {
        float a = 0.f;
        if ((Cse & 3) == 0) {
            const int n4 = Cse >> 2;
            const float4* m4 = reinterpret_cast<const float4*>(m);
            const float4* w4 = reinterpret_cast<const float4*>(Wr + (long long)q * Cse);
#pragma unroll 10
            for (int i = lane; i < n4; i += 32) {
                const float4 x = m4[i], u = __ldg(w4 + i);
                a = fmaf(u.x, x.x, fmaf(u.y, x.y, fmaf(u.z, x.z, fmaf(u.w, x.w, a))));
            }
        } else {
            for (int c = lane; c < Cse; c += 32) a = fmaf(__ldg(Wr + (long long)q * Cse + c), m[c], a);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) r[q] = silu(a + __ldg(br + q));
    }
    __syncthreads();
    // excite: a thread owns channels c, c + 1024 and walks the Sq rows of We^T (Sq, Cse; coalesced over c) once for both
    constexpr int kPer = (kSeMaxC + kSeThreads - 1) / kSeThreads;
    float a[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int c = threadIdx.x + j * kSeThreads;
        a[j] = c < Cse ? __ldg(be + c) : 0.f;
    }
#pragma unroll 8
    for (int q = 0; q < Sq; ++q) {
        const float rq = r[q];
        const float* wq = We + (long long)q * Cse + threadIdx.x;
#pragma unroll
        for (int j = 0; j < kPer; ++j)
            if (threadIdx.x + j * kSeThreads < Cse) a[j] = fmaf(__ldg(wq + j * kSeThreads), rq, a[j]);
    }
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int c = threadIdx.x + j * kSeThreads;
        if (c < C) row[c] = c < Cse ? __fdividef(1.f, 1.f + __expf(-a[j])) : 0.f;
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
    const long long total = (long long)N * Hl * Wl * (C_out >> 3);           // one work item per source cell and channel group
    if ((long long)N * H * W * (C_out >> 3) < (1ll << 31))
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
    memset(&sw, 0, sizeof(sw));
    const float* wf = (const float*)w;
    for (int k = 0; k < 27; ++k)
        for (int co = 0; co < 32; ++co) {
            uint32_t u;
            memcpy(&u, wf + k * 32 + co, 4);
            sw.w[k][co] = (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);      // fp32 -> bf16, round to nearest even
        }
    memcpy(sw.shift, shift, sizeof(sw.shift));
    const long long tiles = (total + 15) / 16;
    stem_conv_kernel<<<grid_for((tiles + kStemWarps - 1) / kStemWarps, 1, 148 * 8), kStemWarps * 32, 0, (cudaStream_t)stream>>>(
        (const float*)img, sw, (__nv_bfloat16*)y, N, H, W, Ho, Wo, pad_h, pad_w);
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
    return launch_dwconv_tma(x, (const float*)w, (const float*)shift, y, (float*)pool, N, H, W, C, Ho, Wo, K, stride, pad_h, pad_w,
                             (cudaStream_t)stream);
}


int mfb_se_fold_bf16(void* pool, float inv_hw, const void* w_reduce, const void* b_reduce, const void* w_expand,
                     const void* b_expand, const void* proj_w, void* out_w, int N, int C, int C_se, int Sq, int Cout, void* stream) {
    if (!pool || !w_reduce || !b_reduce || !w_expand || !b_expand || !proj_w || !out_w) return fail_status(MFB_ERR_INVALID_ARGUMENT, "se_fold: NULL pointer");
    if (N < 1 || Cout < 1 || Sq < 1 || Sq > kSeMaxSq || C < 8 || C > kSeMaxC || (C & 7) || C_se < 1 || C_se > C)
        return fail_status(MFB_ERR_UNSUPPORTED, "se_fold: need C % 8 == 0, C <= 1280, Sq <= 64, C_se <= C");
    if ((uintptr_t)pool & 15) return fail_status(MFB_ERR_INVALID_ARGUMENT, "se_fold: pool must be 16-byte aligned");
    se_mlp_kernel<<<(unsigned)N, kSeThreads, 0, (cudaStream_t)stream>>>((float*)pool, inv_hw, (const float*)w_reduce, (const float*)b_reduce,
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
