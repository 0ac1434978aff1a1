// K4: dense KH x KW convolution (stride 1 or 2, arbitrary low-side zero padding) as an implicit GEMM on the 5th-gen
// tensor cores (tcgen05.mma, accumulator in TMEM, operands staged by TMA), with a fused epilogue.  It carries every
// GEMM-shaped layer of the terrain encoder:
//     Up.conv (two 3x3 conv-BN-GELU)                                   terrain_encoder/lss.py:34-41
//     BevEncode conv1 7x7/2 + ResNet-18 layer1-3 (3x3, 3x3/2, 1x1/2,
//       BN, residual add, ReLU)                                        terrain_encoder/lss.py:104-116,140-151
//     BevEncode heads: 3x3 conv-BN-GELU + 1x1 conv + ScaledTanh/ReLU   terrain_encoder/lss.py:117-139
//     1x1 depthnet                                                     terrain_encoder/lss.py:58
//     EfficientNet-B0 MBConv 1x1 expand (BN+SiLU) / project (BN, skip) terrain_encoder/lss.py:73-94
//
//   t[n,h,w,co] = scale[co] * sum_{dy,dx,ci} x[n, s*h+dy-ph, s*w+dx-pw, ci] * wgt[(n,) co,dy,dx,ci] + shift[co]
//   y = act(t + residual)                                   (bf16 NHWC), or, in head mode,
//   out[n,g,h,w] = head_act_g( sum_{c < BLOCK_N} act(t)[g*BLOCK_N + c] * head_w[g*BLOCK_N + c] + head_b[g] )   (fp32)
//
// Layouts: x, y, residual NHWC bf16; wgt [Cout][KH*KW*Cin] bf16 (K-major), optionally one matrix per image (the
// squeeze-excite scale of an MBConv block folded into its projection weights); scale / shift fp32.
// Tiling: one CTA = 128 output pixels (TH x TW = 8 x 16 patch of one image) x BLOCK_N output channels (64, 128, or 256 for the
// deep-K layers: conv_tcgen05.cu::launch256).
// The K loop runs over (tap, 64-channel chunk): for each, TMA loads the SHIFTED (and for stride 2: element-strided)
// activation patch as a 4-D box {64 ch, TW*s, TH*s, 1} with traversal strides {1,s,s,1} (out-of-image rows/columns are
// zero-filled by the TMA unit == the conv's zero padding; so are the channels beyond Cin when Cin % 64 != 0) and the matching
// weight slab {64, BLOCK_N}; both land in
// 128-byte-swizzled K-major shared tiles, the canonical UMMA operand layout, so no im2col buffer ever exists.
// Persistent CTAs (2 per SM) walk the tiles; two TMEM accumulators overlap one tile's epilogue with the next tile's MMAs.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2-9 = epilogue (tcgen05.ld -> scale/shift (+residual) -> activation -> bf16 16-byte stores | head dot product);
// warp w reads TMEM lane quadrant w % 4 and the column half (w - 2) / 4.  The activation is a template parameter: the
// epilogue of the memory-bound layers is instruction-latency-bound (ncu: 12 cycles per issued instruction with 2 warps per
// scheduler), so per-element branches and libm calls are what it cannot afford.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mfb {
namespace conv {

constexpr int kBlockM = 128;      // output pixels per CTA
constexpr int kTileW = 16;        // patch width  (pixels)
constexpr int kTileH = 8;         // patch height (pixels)
constexpr int kBlockK = 64;       // bf16 channels per K chunk = one 128-byte swizzle row
constexpr int kEpiWarps = 8;      // two per TMEM lane quadrant: each takes half of the tile's columns
constexpr int kThreads = 64 + 32 * kEpiWarps;

enum Act : int { kNone = 0, kRelu = 1, kGelu = 2, kSilu = 3 };
enum HeadAct : int { kHeadNone = 0, kHeadRelu = 1, kHeadScaledTanh = 2 };
constexpr int kMaxHeads = 4;
constexpr int kMaxParamChannels = 1280;   // scale / shift (/ head_w) of every output channel are staged in shared memory once per CTA

struct Params {
    int N, H, W, Cin;              // input  (N,H,W,Cin)
    int Ho, Wo, Cout;              // output (N,Ho,Wo,Cout)
    int KH, KW, stride, pad_h, pad_w, act;
    int tiles_w, tiles_h;
    int per_image_w;               // weights are (N, Cout, KH*KW*Cin): the CTA's image selects the matrix
    __nv_bfloat16* y;              // nullptr in head mode
    const __nv_bfloat16* res;      // residual added before the activation, or nullptr
    const float* scale;
    const float* shift;
    // head mode: a 1x1 convolution to ONE channel per BLOCK_N-channel group + its output activation, fused
    float* head_out;               // (N, Cout / BLOCK_N, Ho, Wo) fp32, or nullptr
    const float* head_w;           // (Cout,)
    float head_b[kMaxHeads], head_lo[kMaxHeads], head_hi[kMaxHeads];
    int head_act[kMaxHeads];
};

constexpr int kXposePitch = 80;   // bytes per staged row: 32 bf16 + 16 B pad (16-byte vector accesses, <= 2-way bank conflicts)

// STAGES = 3 for the deep-K (compute-bound) layers; the memory-bound ones (<= 18 K chunks) run 2 stages and spend the 32 KB on a
// per-warp transpose buffer, so that the epilogue's global stores are 64-byte runs instead of 32 scattered 16-byte pieces
template <int BLOCK_N, int STAGES, bool XPOSE>
struct Smem {
    static constexpr int kABytes = kBlockM * kBlockK * 2;        // 16 KB
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarrierOffset = STAGES * kStageBytes;
    static constexpr int kParamOffset = kBarrierOffset + 256 + 512;           // fp32 scale | shift | head_w of all output channels
    static constexpr int kXposeOffset = kParamOffset + (2 * kMaxParamChannels + kMaxHeads * 128) * 4;
    static constexpr int kTotal = kXposeOffset + (XPOSE ? kEpiWarps * 32 * kXposePitch : 0) + 1024;   // 2 CTAs per SM: <= 113 KB each
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(s_addr(bar)), "r"(parity) : "memory");
    }
}
// Same wait for the two single-thread roles (TMA producer waiting for a free stage, MMA issuer waiting for a drained
// accumulator): between polls the thread sleeps, so that on the memory-bound layers - where both of them wait for the
// epilogue most of the time - their polling does not take issue slots from the eight epilogue warps (ncu, MBConv block-1
// expand: a quarter of all issued instructions were these two loops).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(s_addr(bar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(40);
    }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(s_addr(dst)), "l"(map), "r"(s_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(s_addr(dst)), "l"(map), "r"(s_addr(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(s_addr(dst)), "l"(map), "r"(s_addr(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s_addr(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(s_addr(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, SM100 version 1):
// start address >> 4 | LBO = 1 (unused for swizzled K-major) | SBO = 1024 B (8 rows x 128 B) | SWIZZLE_128B
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(const void* smem_tile) {
    const uint64_t addr = (uint64_t)(s_addr(smem_tile) >> 4) & 0x3FFFull;
    return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 output rounding): 2 SFU ops + 10 FMA-class
// instructions, no branches (libm's erff is ~3x that with a data-dependent branch)
__device__ __forceinline__ float erf_fast(float x) {
    const float ax = fabsf(x);
    const float t = rcp_approx(fmaf(0.3275911f, ax, 1.f));
    float poly = fmaf(t, 1.061405429f, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    const float e = ex2_approx(ax * ax * -1.4426950408889634f);
    return copysignf(fmaf(-poly, e, 1.f), x);
}

// Activation of u = kActPre<ACT> * (acc * scale + shift): for SiLU the epilogue works on h = x / 2 (the factor is folded into
// the staged scale / shift), x sigmoid(x) = h + h tanh(h): one FMA + one SFU op per element instead of FMUL + FMA + SFU.
template <int ACT> constexpr float kActPre = ACT == kSilu ? 0.5f : 1.f;
template <int ACT>
__device__ __forceinline__ float apply_act(float u) {
    if (ACT == kRelu) return fmaxf(u, 0.f);
    if (ACT == kGelu) return 0.5f * u * (1.f + erf_fast(u * 0.70710678118654752f));      // nn.GELU() (erf form)
    if (ACT == kSilu) return fmaf(u, tanh_approx(u), u);
    return u;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s_addr(bar)) : "memory");
}

// Persistent: grid = min(tiles, 2 CTAs per SM); every CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... with
// (output-channel block) fastest, so the CTAs that share an activation patch run at the same time and hit it in L2.
// Three pipelines (guide: "canonical Blackwell GEMM"): smem full/empty (TMA <-> MMA, kStages deep, runs across tile
// boundaries so the next tile's operands stream in during this tile's epilogue), TMEM full/empty (MMA <-> epilogue, two
// accumulators of BLOCK_N columns: the MMAs of tile i+1 overlap the epilogue of tile i), and the tile walk itself.
template <int BLOCK_N, int ACT, bool HEAD, int kStages, bool XPOSE, bool RES>
__global__ void __launch_bounds__(kThreads, 2)
conv_bn_act_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B needs 1024-B alignment.  The offset is added to the __shared__ array itself (not to an integer cast of it), so
    // that the compiler keeps the address space and emits LDS / STS instead of generic LD / ST for the epilogue's accesses.
    uint8_t* smem = smem_raw + ((1024u - (s_addr(smem_raw) & 1023u)) & 1023u);
    using S = Smem<BLOCK_N, kStages, XPOSE>;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarrierOffset);
    uint64_t* empty = full + kStages;
    uint64_t* acc_full = empty + kStages;       // [2]
    uint64_t* acc_empty = acc_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_blocks = (p.Cout + BLOCK_N - 1) / BLOCK_N;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    const unsigned total = (unsigned)tiles_per_img * p.N * n_blocks;     // host guarantees < 2^31: 32-bit div / mod on the tile walk
    // Cin need not be a multiple of 64: the last chunk's channels beyond Cin are outside the tensor map's extent, which the
    // TMA unit fills with zeros (for the weights too, whose rows are KH*KW*Cin long), so they add nothing to the sum
    const int chunks_per_tap = (p.Cin + kBlockK - 1) / kBlockK;
    const int k_iters = p.KH * p.KW * chunks_per_tap;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_w) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, kEpiWarps); }     // every epilogue warp drains its part
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * BLOCK_N);      // two fp32 accumulators of BLOCK_N columns (power of two >= 32)
    // Epilogue parameters: a persistent CTA sees every output-channel block many times, and ncu showed the epilogue waiting on
    // exactly these loads (long_scoreboard 11.6 cycles per issued instruction, L1 hit rate 19 %: the streaming stores leave
    // nothing of them in L1).  Staged once; padded to the tile grid with (1, 0, 0) so that no read needs a bounds test.
    float* s_scale = reinterpret_cast<float*>(smem + S::kParamOffset);
    float* s_shift = s_scale + kMaxParamChannels;
    float* s_head = s_shift + kMaxParamChannels;
    for (int i = threadIdx.x; i < n_blocks * BLOCK_N; i += kThreads) {
        const bool in = i < p.Cout;
        s_scale[i] = in ? kActPre<ACT> * __ldg(p.scale + i) : 1.f;
        s_shift[i] = in ? kActPre<ACT> * __ldg(p.shift + i) : 0.f;
        if (HEAD) s_head[i] = in ? __ldg(p.head_w + i) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            uint32_t it = 0;
            for (unsigned t = blockIdx.x; t < total; t += gridDim.x) {
                const unsigned pt = t / (unsigned)n_blocks;
                const int nb = (int)(t - pt * n_blocks);
                const unsigned n_ = pt / (unsigned)tiles_per_img, rem = pt - n_ * tiles_per_img;
                const int th = (int)(rem / (unsigned)p.tiles_w), tw = (int)(rem - th * p.tiles_w), n = (int)n_;
                const int w0 = tw * kTileW, h0 = th * kTileH, n0 = nb * BLOCK_N;
                for (int i = 0; i < k_iters; ++i, ++it) {
                    const int s = it % kStages;
                    mbar_wait_relaxed(empty + s, ((it / kStages) & 1) ^ 1);
                    const int tap = i / chunks_per_tap, c0 = (i - tap * chunks_per_tap) * kBlockK;
                    const int dy = tap / p.KW, dx = tap - dy * p.KW;
                    uint8_t* a_dst = smem + s * S::kStageBytes;
                    uint8_t* b_dst = a_dst + S::kABytes;
                    mbar_expect_tx(full + s, S::kStageBytes);
                    // input pixel of output (h, w) under tap (dy, dx): (stride*h + dy - pad_h, stride*w + dx - pad_w); the
                    // tensor map traverses W and H with step `stride`, so the box lands as TH x TW consecutive rows
                    tma_load_4d(a_dst, &tmap_x, full + s, c0, w0 * p.stride + dx - p.pad_w, h0 * p.stride + dy - p.pad_h, n);
                    if (p.per_image_w) tma_load_3d(b_dst, &tmap_w, full + s, tap * p.Cin + c0, n0, n);
                    else tma_load_2d(b_dst, &tmap_w, full + s, tap * p.Cin + c0, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer (one thread) =====
            constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N);
            uint32_t it = 0, lt = 0;
            for (unsigned t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
                const uint32_t as = lt & 1;
                mbar_wait_relaxed(acc_empty + as, ((lt >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + as * BLOCK_N;
                for (int i = 0; i < k_iters; ++i, ++it) {
                    const int s = it % kStages;
                    mbar_wait(full + s, (it / kStages) & 1);
                    tc_fence_after();
                    const uint8_t* a_src = smem + s * S::kStageBytes;
                    const uint64_t a_desc = make_kmajor_sw128_desc(a_src);
                    const uint64_t b_desc = make_kmajor_sw128_desc(a_src + S::kABytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (>>4) address field
                        umma_bf16(tmem_acc, a_desc + 2 * k, b_desc + 2 * k, idesc, (i | k) != 0);
                    }
                    umma_commit(empty + s);                      // frees the stage once these MMAs have read it
                }
                umma_commit(acc_full + as);                      // accumulator complete
            }
        }
    } else {
        // ===== epilogue: warp w owns TMEM lane quadrant w % 4 and column half (w - 2) / 4 =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        constexpr int kColsPerWarp = BLOCK_N / 2;
        const int c_lo = half * kColsPerWarp;
        const int m = q * 32 + lane;                         // accumulator row == pixel inside the patch
        const int hh = m / kTileW, ww = m - hh * kTileW;
        float* head_part = reinterpret_cast<float*>(tmem_slot + 4);     // [kBlockM] partial head dots of the upper column half
        uint8_t* const xbuf = smem + S::kXposeOffset + (warp - 2) * 32 * kXposePitch;   // this warp's transpose buffer (XPOSE)
        uint32_t lt = 0;
        for (unsigned t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
            const unsigned pt = t / (unsigned)n_blocks;
            const int nb = (int)(t - pt * n_blocks);
            const unsigned n_ = pt / (unsigned)tiles_per_img, rem = pt - n_ * tiles_per_img;
            const int th = (int)(rem / (unsigned)p.tiles_w), tw = (int)(rem - th * p.tiles_w), n = (int)n_;
            const int n0 = nb * BLOCK_N;
            const int h = th * kTileH + hh, w = tw * kTileW + ww;
            const bool in_image = (h < p.Ho) && (w < p.Wo);
            const long long pix = ((long long)n * p.Ho + h) * p.Wo + w;
            const uint32_t as = lt & 1;
            mbar_wait(acc_full + as, (lt >> 1) & 1);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + as * BLOCK_N + ((uint32_t)(q * 32) << 16);
            if (HEAD) {
                // fused 1x1 head: one output channel per BLOCK_N-channel group (this tile's group = nb)
                float dot = 0.f;
#pragma unroll 1
                for (int c = c_lo; c < c_lo + kColsPerWarp; c += 32) {
                    uint32_t r[32];
                    tmem_ld32(tmem_acc + c, r);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int co = n0 + c + j;
                        const float v = apply_act<ACT>(fmaf(__uint_as_float(r[j]), s_scale[co], s_shift[co]));
                        dot = fmaf(v, s_head[co], dot);
                    }
                }
                // the two warps of a lane quadrant each hold half of the dot product: combine through shared memory
                if (half == 1) head_part[m] = dot;
                asm volatile("bar.sync %0, 64;" :: "r"(1 + q) : "memory");
                if (half == 0 && in_image) {
                    float o = dot + head_part[m] + p.head_b[nb];
                    if (p.head_act[nb] == kHeadRelu) o = fmaxf(o, 0.f);
                    else if (p.head_act[nb] == kHeadScaledTanh) o = p.head_lo[nb] + (p.head_hi[nb] - p.head_lo[nb]) * (tanhf(o) + 1.f) * 0.5f;
                    p.head_out[((long long)n * n_blocks + nb) * p.Ho * p.Wo + (long long)h * p.Wo + w] = o;
                }
                asm volatile("bar.sync %0, 64;" :: "r"(1 + q) : "memory");      // head_part is free for the next tile
            } else {
                __nv_bfloat16* out = p.y + pix * p.Cout + n0;
                // transposed write-out: output pixel of the four accumulator rows this lane stores (-1: outside the image); per tile,
                // not per 32-column chunk (the 64-bit index arithmetic was ~1.5 instructions per output element)
                int xrow_pix[4];
                if (XPOSE) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int mm = q * 32 + 8 * k + (lane >> 2);
                        const int h2 = th * kTileH + mm / kTileW, w2 = tw * kTileW + (mm % kTileW);
                        xrow_pix[k] = (h2 < p.Ho && w2 < p.Wo) ? (n * p.Ho + h2) * p.Wo + w2 : -1;
                    }
                }
                // RES is a template parameter: as a run-time test the residual add was compiled into ~64 predicated instructions
                // per 32 columns that took issue slots in every layer without a residual (the epilogue is issue-bound)
                const __nv_bfloat16* rsd = RES ? p.res + pix * p.Cout + n0 : nullptr;
#pragma unroll 1
                for (int c = c_lo; c < c_lo + kColsPerWarp; c += 32) {
                    if (n0 + c >= p.Cout) break;              // warp-uniform: nothing but padding columns left
                    uint32_t r[32];
                    tmem_ld32(tmem_acc + c, r);
                    if (in_image) {
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            const int co = n0 + c + 8 * g8;
                            if (co < p.Cout) {                    // Cout is a multiple of 8, not necessarily of BLOCK_N
                                const float4 sc0 = *reinterpret_cast<const float4*>(s_scale + co), sc1 = *reinterpret_cast<const float4*>(s_scale + co + 4);
                                const float4 sh0 = *reinterpret_cast<const float4*>(s_shift + co), sh1 = *reinterpret_cast<const float4*>(s_shift + co + 4);
                                const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
                                const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
                                float v[8];
#pragma unroll
                                for (int j = 0; j < 8; j += 2) {           // packed f32x2: the accumulator columns arrive as register pairs
                                    const float2 a2 = make_float2(__uint_as_float(r[8 * g8 + j]), __uint_as_float(r[8 * g8 + j + 1]));
                                    const float2 v2 = __ffma2_rn(a2, make_float2(sc[j], sc[j + 1]), make_float2(sh[j], sh[j + 1]));
                                    v[j] = v2.x; v[j + 1] = v2.y;
                                }
                                if (RES) {
                                    const uint4 rv = __ldg(reinterpret_cast<const uint4*>(rsd + c + 8 * g8));
                                    const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const float2 f = __bfloat1622float2(r2[j]);
                                        v[2 * j] = fmaf(kActPre<ACT>, f.x, v[2 * j]); v[2 * j + 1] = fmaf(kActPre<ACT>, f.y, v[2 * j + 1]);
                                    }
                                }
                                uint32_t pk[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float o0, o1;
                                    if (ACT == kSilu) {
                                        const float2 u2 = make_float2(v[2 * j], v[2 * j + 1]);
                                        const float2 o2 = __ffma2_rn(u2, make_float2(tanh_approx(u2.x), tanh_approx(u2.y)), u2);
                                        o0 = o2.x; o1 = o2.y;
                                    } else {
                                        o0 = apply_act<ACT>(v[2 * j]); o1 = apply_act<ACT>(v[2 * j + 1]);
                                    }
                                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(o0, o1);
                                    pk[j] = *reinterpret_cast<const uint32_t*>(&b2);
                                }
                                if (XPOSE) *reinterpret_cast<uint4*>(xbuf + lane * kXposePitch + g8 * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                else *reinterpret_cast<uint4*>(out + c + 8 * g8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            }
                        }
                    }
                    if (XPOSE) {
                        // transposed write-out: 4 lanes cover the 64 bytes (32 channels) of one pixel, 8 pixels per instruction
                        __syncwarp();
                        const int part = lane & 3;
                        const int co = n0 + c + 8 * part;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int row = 8 * k + (lane >> 2);             // row of this warp's 32 accumulator rows
                            if (xrow_pix[k] >= 0 && co < p.Cout) {
                                const uint4 val = *reinterpret_cast<const uint4*>(xbuf + row * kXposePitch + part * 16);
                                *reinterpret_cast<uint4*>(p.y + (long long)xrow_pix[k] * p.Cout + co) = val;
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            // this warp has read everything it needs from the accumulator: hand it back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + as);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * BLOCK_N);
}

}  // namespace conv
}  // namespace mfb
