// K7d: depthwise KxK convolution + folded BatchNorm + swish (+ squeeze-excite spatial sums), NHWC bf16, fed by TMA.
// Replaces nn.Conv2d(groups=C) -> BatchNorm2d -> MemoryEfficientSwish of an EfficientNet MBConv block (efficientnet_pytorch
// model.py MBConvBlock.forward, called from the reference's CamEncode, lss.py:73-94) and the adaptive_avg_pool2d that feeds its
// squeeze-excite branch.
//
// Shape of the kernel (measured on B200 at the EfficientNet-B0 layer shapes, tools/dw_bench.py; the cp.async tile / strip kernels
// it supersedes ran the 16 layers of a 64-image 512x512 batch in 2.40 ms, this one in 1.29 ms):
//   * a CTA owns an output tile x one channel slab (64 channels = 128-byte rows; 32 channels when 64 would idle half of the
//     lanes: C <= 32).  ONE cp.async.bulk.tensor per tile (4-D box; out-of-bounds = the convolution's zero padding,
//     ragged channel counts zero-filled) brings the whole input patch, double-buffered along a persistent walk over contiguous
//     (image, slab, tile) ranges: no staging instructions, no bounds arithmetic;
//   * a thread owns 2 channels x (4 x 4) output pixels (stride 2: 2 x 4): an input pixel is unpacked once (one shift, one mask)
//     and used by up to K*K outputs from registers; all math is packed f32x2 over the channel pair (FFMA2 / FMUL2 / FADD2);
//   * the lane's BN-folded taps and shift live in registers (2 K^2 + 2 values), reloaded only when the walk changes slab;
//   * squeeze-excite sums stay in registers until the walk leaves the (image, slab).
// Bound: the 5x5 layers run at 12-14 TFMA/s, a third of the fp32 peak: ncu shows the FMA pipe 48 % busy (an FFMA2 holds it for two
// cycles) and 51 % of the issue slots used - a scalar-FFMA build ran in exactly the same time, so it is the phase structure of a
// thread (load + unpack phases idle the pipe, FMA-dense phases saturate it) with 16 warps per SM, not the instruction count.  The
// 3x3 / stride-2 layers are bound by their loads (2.3-4.5 TB/s).  Tensor cores (a banded-Toeplitz formulation) are the next step.
#include <cstdint>
#include <mutex>
#include <string>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/monoforce_b200.h"

namespace mfb {
void count_launch();
int fail_status(int code, const std::string& msg);

namespace enc {

constexpr int kThreads = 256;

// LANES threads (2 channels each) cover one pixel block: LANES = 32 -> 64-channel slabs (128-byte TMA rows, one warp per block),
// LANES = 16 -> 32-channel slabs for the layers with few channels (two blocks per warp).
template <int K, int S, int LANES, int PASSES, int STAGES>
struct DwCfg {
    static constexpr int kSlab = 2 * LANES;                   // channels per CTA tile
    static constexpr int kBlocks = kThreads / LANES;          // pixel blocks per CTA
    static constexpr int BH = S == 1 ? 4 : 2, BW = 4;         // output pixels per thread
    static constexpr int NBW = 4, NBH = kBlocks / NBW;        // blocks per pass: 2 x 4 or 4 x 4
    // one TMA load feeds kPasses row bands of the tile (stride 1, 64-channel slabs: 16 x 16 outputs in two 8-row bands):
    // half as many loads / barriers / index updates per output and a smaller halo
    static constexpr int kPasses = PASSES;
    static constexpr int OTH = NBH * BH * kPasses, OTW = NBW * BW;      // output tile
    static constexpr int IBH = (BH - 1) * S + K, IBW = (BW - 1) * S + K;     // input block of a thread
    static constexpr int ITH = (OTH - 1) * S + K, ITW = (OTW - 1) * S + K;   // input patch of the CTA
    static constexpr int kPitch = kSlab * 2;                  // bytes per staged pixel
    static constexpr int kBufBytes = ITH * ITW * kPitch;
    static constexpr int kBufStride = (kBufBytes + 127) / 128 * 128;
    static constexpr int kStages = STAGES;
    static constexpr int kPoolOffset = kStages * kBufStride;
    static constexpr int kBarOffset = kPoolOffset + kSlab * 4;
    static constexpr int kTotal = kBarOffset + 64 + 128;      // + slack for the 128-byte alignment of the base
};

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// swish of a channel pair, x sigmoid(x) = x/2 (1 + tanh(x/2)): two packed FMA-class instructions + two SFU ops (tanh.approx:
// 2^-11 relative, below bf16 rounding)
__device__ __forceinline__ float2 silu2(float2 v) {
    const float2 h = __fmul2_rn(v, make_float2(0.5f, 0.5f));
    float2 t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
    return __ffma2_rn(h, t, h);
}

template <int K, int S, int LANES, int PASSES, int STAGES>
__global__ void __launch_bounds__(kThreads, 2)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ w, const float* __restrict__ shift,
                  __nv_bfloat16* __restrict__ y, float* __restrict__ pool, int N, int C, int Ho, int Wo, int ph, int pw) {
    using Cfg = DwCfg<K, S, LANES, PASSES, STAGES>;
    constexpr int BH = Cfg::BH, BW = Cfg::BW, IBH = Cfg::IBH, IBW = Cfg::IBW, ITW = Cfg::ITW, kSlab = Cfg::kSlab;
    extern __shared__ unsigned char dw_raw[];
    const uint32_t base = (s_addr(dw_raw) + 127u) & ~127u;
    float* const pool_s = reinterpret_cast<float*>(dw_raw + (base - s_addr(dw_raw)) + Cfg::kPoolOffset);
    const uint32_t bar0 = base + Cfg::kBarOffset;

    const int tid = threadIdx.x;
    const int cl = tid % LANES;                    // channel lane: channels 2 cl, 2 cl + 1 of the slab
    const int blk = tid / LANES, bh = blk / Cfg::NBW, bw = blk % Cfg::NBW;

    const int tiles_w = (Wo + Cfg::OTW - 1) / Cfg::OTW, tiles_h = (Ho + Cfg::OTH - 1) / Cfg::OTH;
    const int sp = tiles_w * tiles_h;
    const int slabs = (C + kSlab - 1) / kSlab;
    const long long total = (long long)N * slabs * sp;
    const long long lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;
    if (lo >= hi) return;

    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < STAGES; ++k) mbar_init(bar0 + 8 * k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < kSlab) pool_s[tid] = 0.f;
    __syncthreads();

    // walk over contiguous tasks (n, slab, th, tw), tw fastest: decoded once, then advanced incrementally
    struct Task { int n, slab, th, tw; };
    auto advance = [&](Task& k) {
        if (++k.tw == tiles_w) {
            k.tw = 0;
            if (++k.th == tiles_h) {
                k.th = 0;
                if (++k.slab == slabs) { k.slab = 0; ++k.n; }
            }
        }
    };
    Task cur;
    {
        const int key = (int)(lo / sp), q = (int)(lo - (long long)key * sp);
        cur.n = key / slabs; cur.slab = key - cur.n * slabs;
        cur.th = q / tiles_w; cur.tw = q - cur.th * tiles_w;
    }
    auto issue = [&](const Task& k, int s) {        // thread 0: bring the input patch of a task into buffer s
        mbar_expect_tx(bar0 + 8 * s, Cfg::kBufBytes);
        tma_load_4d(base + s * Cfg::kBufStride, &tmap_x, bar0 + 8 * s, k.slab * kSlab, k.tw * Cfg::OTW * S - pw, k.th * Cfg::OTH * S - ph,
                    k.n);
    };
    // prologue: the first STAGES - 1 loads; `ahead` = the task STAGES - 1 positions after `cur`
    const int n_tasks = (int)(hi - lo);
    Task ahead = cur;
    for (int k = 0; k < STAGES - 1; ++k) {
        if (k < n_tasks && tid == 0) issue(ahead, k);
        advance(ahead);
    }

    float2 psum = make_float2(0.f, 0.f);
    float2 wreg[K * K], sh2 = make_float2(0.f, 0.f);   // this lane's two channels of the BN-folded taps and shift
    int cur_key = -1, cur_slab = -1;

    auto flush_pool = [&](int key) {                 // block-wide: all threads call it
        const int n = key / slabs, c0 = (key - n * slabs) * kSlab;
        if (LANES == 16) {
            psum.x += __shfl_xor_sync(0xffffffffu, psum.x, 16);
            psum.y += __shfl_xor_sync(0xffffffffu, psum.y, 16);
        }
        if ((tid & 31) < LANES) {
            atomicAdd(pool_s + cl * 2, psum.x);
            atomicAdd(pool_s + cl * 2 + 1, psum.y);
        }
        __syncthreads();
        if (tid < kSlab) {
            const float v = pool_s[tid];
            if (c0 + tid < C && v != 0.f) atomicAdd(pool + (long long)n * C + c0 + tid, v);
            pool_s[tid] = 0.f;
        }
        psum = make_float2(0.f, 0.f);
    };

    for (int i = 0; i < n_tasks; ++i) {
        const int s = i % STAGES;
        // the buffer of task i + STAGES - 1 was released by the barrier that ended iteration i - 1
        if (i + STAGES - 1 < n_tasks && tid == 0) issue(ahead, (i + STAGES - 1) % STAGES);
        advance(ahead);
        const int n = cur.n, slab = cur.slab, key = n * slabs + slab;
        const int c = slab * kSlab + cl * 2;                     // first of this lane's two channels (C is even)
        if (key != cur_key) {
            if (pool && cur_key >= 0) flush_pool(cur_key);       // (its barrier also orders pool_s against the next flush)
            if (slab != cur_slab) {
#pragma unroll
                for (int tap = 0; tap < K * K; ++tap)
                    wreg[tap] = c < C ? __ldg(reinterpret_cast<const float2*>(w + (long long)tap * C + c)) : make_float2(0.f, 0.f);
                sh2 = c < C ? __ldg(reinterpret_cast<const float2*>(shift + c)) : make_float2(0.f, 0.f);
                cur_slab = slab;
            }
            cur_key = key;
        }
        mbar_wait(bar0 + 8 * s, (uint32_t)(i / STAGES) & 1u);

#pragma unroll 1
        for (int pass = 0; pass < Cfg::kPasses; ++pass) {
        const int hb = (pass * Cfg::NBH + bh) * BH;                                  // row of this thread's block inside the tile
        const int h0 = cur.th * Cfg::OTH + hb, w0 = cur.tw * Cfg::OTW + bw * BW;     // first output pixel of this thread's block
        if (h0 < Ho && w0 < Wo && c < C) {
            const uint32_t in0 = base + s * Cfg::kBufStride + ((hb * S) * ITW + bw * BW * S) * Cfg::kPitch + cl * 4;
            float2 acc[BH][BW];
#pragma unroll
            for (int a = 0; a < BH; ++a)
#pragma unroll
                for (int b = 0; b < BW; ++b) acc[a][b] = sh2;
#pragma unroll
            for (int r = 0; r < IBH; ++r) {
                float2 tv[IBW];
#pragma unroll
                for (int e = 0; e < IBW; ++e) {
                    const uint32_t u = lds32(in0 + (r * ITW + e) * Cfg::kPitch);
                    tv[e] = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
                }
#pragma unroll
                for (int a = 0; a < BH; ++a) {
                    const int dy = r - a * S;
                    if (dy >= 0 && dy < K) {
#pragma unroll
                        for (int dx = 0; dx < K; ++dx)
#pragma unroll
                            for (int b = 0; b < BW; ++b) acc[a][b] = __ffma2_rn(tv[b * S + dx], wreg[dy * K + dx], acc[a][b]);
                    }
                }
            }
            __nv_bfloat16* yrow = y + (((long long)n * Ho + h0) * Wo + w0) * C + c;
#pragma unroll
            for (int a = 0; a < BH; ++a) {
                if (h0 + a < Ho) {
#pragma unroll
                    for (int b = 0; b < BW; ++b) {
                        if (w0 + b < Wo) {
                            const float2 v = silu2(acc[a][b]);
                            *reinterpret_cast<__nv_bfloat162*>(yrow + b * C) = __floats2bfloat162_rn(v.x, v.y);
                            psum = __fadd2_rn(psum, v);
                        }
                    }
                }
                yrow += (long long)Wo * C;
            }
        }
        }
        advance(cur);
        __syncthreads();                 // everyone is done with buffer s: it is the TMA target of a later iteration
    }
    if (pool) flush_pool(cur_key);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time dependency on libcuda
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

template <int K, int S, int LANES, int PASSES, int STAGES>
static int launch_ks(const void* x, const float* w, const float* shift, void* y, float* pool, int N, int H, int W, int C, int Ho,
                     int Wo, int pad_h, int pad_w, cudaStream_t st) {
    using Cfg = DwCfg<K, S, LANES, PASSES, STAGES>;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail_status(MFB_ERR_CUDA, "dwconv: cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap mx;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)Cfg::kSlab, (cuuint32_t)Cfg::ITW, (cuuint32_t)Cfg::ITH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    // L2 promotion 64 B: a slab row is 64 or 128 bytes of a pixel whose other channels belong to other CTAs; with 128-byte promotion
    // every row pulled its neighbours' bytes from DRAM too (ncu, 144 channels: 760 MB read for a 302 MB input; 260 -> 227 us)
    const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail_status(MFB_ERR_CUDA, "dwconv: cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    const long long total = (long long)N * ((C + Cfg::kSlab - 1) / Cfg::kSlab) * ((Ho + Cfg::OTH - 1) / Cfg::OTH) *
                            ((Wo + Cfg::OTW - 1) / Cfg::OTW);
    static int sms = 0;
    if (!sms) {
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        sms = n;
    }
    const long long resident = (Cfg::kTotal > 113 * 1024 ? 1ll : 2ll) * sms;
    auto kern = dwconv_tma_kernel<K, S, LANES, PASSES, STAGES>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kTotal) != cudaSuccess)
        return fail_status(MFB_ERR_CUDA, "dwconv: cudaFuncSetAttribute failed");
    kern<<<(unsigned)(total < resident ? total : resident), kThreads, Cfg::kTotal, st>>>(mx, w, shift, (__nv_bfloat16*)y, pool, N, C, Ho,
                                                                                        Wo, pad_h, pad_w);
    count_launch();
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_status(MFB_ERR_CUDA, std::string("dwconv (tma) launch: ") + cudaGetErrorString(e));
    return MFB_OK;
}

// x (N,H,W,C) bf16 -> y (N,Ho,Wo,C) bf16; w (K*K, C) fp32 BN-folded, shift (C,) fp32, pool (N,C) fp32 += spatial sums or nullptr
int launch_dwconv_tma(const void* x, const float* w, const float* shift, void* y, float* pool, int N, int H, int W, int C, int Ho,
                      int Wo, int K, int stride, int pad_h, int pad_w, cudaStream_t st) {
    // 32-channel slabs only when 64-channel ones would idle half of the lanes (C <= 32).  Measured per layer on B200 with both
    // widths (tools/dw_bench.py): at C = 96 and 144 the 128-byte rows of the wide slabs beat the better lane use of the narrow
    // ones (config 4: 221 -> 203 us, 227 -> 189 us, 121 -> 115 us), at C = 32 the narrow ones win 129 vs 227 us.
    const bool narrow = ((C + 63) / 64) * 64 >= 2 * ((C + 31) / 32) * 32;
#define MFB_DW_ARGS x, w, shift, y, pool, N, H, W, C, Ho, Wo, pad_h, pad_w, st
#define MFB_DW_CASE(KK, SS)                                                                \
    if (K == KK && stride == SS)                                                           \
        return narrow ? launch_ks<KK, SS, 16, 1, 2>(MFB_DW_ARGS) : launch_ks<KK, SS, 32, (SS == 1 ? 2 : 1), 2>(MFB_DW_ARGS);
    MFB_DW_CASE(3, 1)
    MFB_DW_CASE(5, 1)
    MFB_DW_CASE(3, 2)
    MFB_DW_CASE(5, 2)
#undef MFB_DW_ARGS
#undef MFB_DW_CASE
    return fail_status(MFB_ERR_UNSUPPORTED, "dwconv: kernel size must be 3 or 5, stride 1 or 2");
}

}  // namespace enc
}  // namespace mfb
