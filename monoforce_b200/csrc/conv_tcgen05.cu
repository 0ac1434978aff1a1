// Host side of K4 (conv_tcgen05.cuh): TMA tensor maps + launch, exported through the C ABI.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/monoforce_b200.h"
#include "conv_tcgen05.cuh"

namespace mfb {
void count_launch();
int fail_status(int code, const std::string& msg);

namespace conv {

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

static int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

template <int BLOCK_N, int ACT, bool HEAD, int STAGES, bool XPOSE, bool RES>
static int launch_v(const CUtensorMap& mx, const CUtensorMap& mw, const Params& p, cudaStream_t st) {
    auto kern = conv_bn_act_kernel<BLOCK_N, ACT, HEAD, STAGES, XPOSE, RES>;
    const int smem = Smem<BLOCK_N, STAGES, XPOSE>::kTotal;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return fail_status(MFB_ERR_CUDA, "conv: cudaFuncSetAttribute failed");
    const long long tiles = (long long)p.tiles_w * p.tiles_h * p.N * ((p.Cout + BLOCK_N - 1) / BLOCK_N);
    if (tiles >= (1ll << 31)) return fail_status(MFB_ERR_UNSUPPORTED, "conv: more than 2^31 tiles");
    const long long resident = (BLOCK_N == 256 ? 1ll : 2ll) * sm_count();     // persistent: 2 CTAs per SM walk the tiles (1 for 256-wide tiles)
    dim3 grid((unsigned)(tiles < resident ? tiles : resident));
    kern<<<grid, kThreads, smem, st>>>(mx, mw, p);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_status(MFB_ERR_CUDA, std::string("conv launch: ") + cudaGetErrorString(e));
    return MFB_OK;
}

// memory-bound layers (few K chunks): 2 stages + transposed (coalesced) stores; deep-K layers: 3 stages, direct stores
template <int BLOCK_N, int ACT, bool HEAD>
static int launch(const CUtensorMap& mx, const CUtensorMap& mw, const Params& p, cudaStream_t st) {
    const int k_iters = p.KH * p.KW * ((p.Cin + kBlockK - 1) / kBlockK);
    // the residual epilogue is a separate instantiation (in the network: MBConv project + identity, ResNet BasicBlock + ReLU)
    constexpr bool kResOk = !HEAD;
    if (kResOk && p.res) {
        if (k_iters <= 18) return launch_v<BLOCK_N, ACT, false, 2, true, kResOk>(mx, mw, p, st);
        return launch_v<BLOCK_N, ACT, false, 3, false, kResOk>(mx, mw, p, st);
    }
    if (!HEAD && k_iters <= 18) return launch_v<BLOCK_N, ACT, false, 2, true, false>(mx, mw, p, st);
    return launch_v<BLOCK_N, ACT, HEAD, 3, false, false>(mx, mw, p, st);
}

// Deep-K layers whose output channel count is a multiple of 256 (camera / BEV Up convolutions, ResNet layer3): 128 x 256 tiles, one
// CTA per SM.  At 128 x 128 the tensor pipe waits for shared-memory fills (ncu, camera Up.conv[0]: pipe 50 % busy, the MMA issuer
// and the epilogue parked on the full / accumulator barriers, L2 hit 95 %): every k-iteration moves 16 KB of activations + 16 KB
// of weights per 2.1 MFLOP; the wide tile moves 16 + 32 KB per 4.2 MFLOP, a quarter less L2 -> shared-memory traffic per flop.
template <int ACT>
static int launch256(const CUtensorMap& mx, const CUtensorMap& mw, const Params& p, cudaStream_t st) {
    if (p.res) return launch_v<256, ACT, false, 3, false, true>(mx, mw, p, st);
    return launch_v<256, ACT, false, 3, false, false>(mx, mw, p, st);
}

}  // namespace conv
}  // namespace mfb

using namespace mfb;
using namespace mfb::conv;

extern "C" int mfb_conv2d_bf16(const mfb_conv_desc* d, const void* x, const void* wgt, const void* scale, const void* shift,
                               const void* residual, void* y, const void* head_w, void* head_out, void* stream) {
    if (!d) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: desc is NULL");
    if (!x || !wgt || !scale || !shift) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: NULL pointer");
    const bool head = d->n_heads > 0;
    if (head ? (!head_w || !head_out) : !y) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: output pointer is NULL");
    if (d->N < 1 || d->H < 1 || d->W < 1 || d->Ho < 1 || d->Wo < 1) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: sizes must be positive");
    if ((long long)d->N * d->Ho * d->Wo >= (1ll << 31)) return fail_status(MFB_ERR_UNSUPPORTED, "conv: more than 2^31 output pixels");
    if (d->KH < 1 || d->KH > 7 || d->KW < 1 || d->KW > 7) return fail_status(MFB_ERR_UNSUPPORTED, "conv: kernel size must be in [1, 7]");
    if (d->stride != 1 && d->stride != 2) return fail_status(MFB_ERR_UNSUPPORTED, "conv: stride must be 1 or 2");
    if (d->pad_h < 0 || d->pad_w < 0 || d->pad_h >= d->KH || d->pad_w >= d->KW) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: padding must be in [0, K)");
    // every output pixel's window must start inside the zero-padded image
    if ((d->Ho - 1) * d->stride - d->pad_h >= d->H || (d->Wo - 1) * d->stride - d->pad_w >= d->W)
        return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: output size does not fit the input / stride / padding");
    if (d->Cin < 8 || d->Cin % 8) return fail_status(MFB_ERR_UNSUPPORTED, "conv: Cin must be a multiple of 8 (16-byte rows for TMA)");
    if (d->Cout < 8 || d->Cout % 8) return fail_status(MFB_ERR_UNSUPPORTED, "conv: Cout must be a multiple of 8");
    if (d->act < kNone || d->act > kSilu) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: unknown activation");
    if (d->Cout > kMaxParamChannels - 128) return fail_status(MFB_ERR_UNSUPPORTED, "conv: at most 1152 output channels");
    if (((uintptr_t)x | (uintptr_t)wgt | (uintptr_t)y | (uintptr_t)residual) & 15) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: tensors must be 16-byte aligned");
    const int k_chunks = d->KH * d->KW * ((d->Cin + kBlockK - 1) / kBlockK);
    const long long tiles256 = (long long)((d->Wo + kTileW - 1) / kTileW) * ((d->Ho + kTileH - 1) / kTileH) * d->N * (d->Cout / 256);
    const bool wide_n = !head && k_chunks > 18 && d->Cout % 256 == 0 && tiles256 >= sm_count();   // enough tiles for one CTA per SM
    const int block_n = wide_n ? 256 : (d->Cout > 64 ? 128 : 64);
    if (((uintptr_t)scale | (uintptr_t)shift) & 15) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: scale / shift must be 16-byte aligned");
    if (head) {
        if (d->n_heads > kMaxHeads || d->n_heads * block_n != d->Cout)
            return fail_status(MFB_ERR_UNSUPPORTED, "conv: head mode needs Cout == n_heads * 128 (or 64) and n_heads <= 4");
        if (residual) return fail_status(MFB_ERR_UNSUPPORTED, "conv: head mode takes no residual");
        for (int i = 0; i < d->n_heads; ++i)
            if (d->head_act[i] < kHeadNone || d->head_act[i] > kHeadScaledTanh) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: unknown head activation");
    }
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail_status(MFB_ERR_CUDA, "conv: cuTensorMapEncodeTiled is not available from the driver");

    const int N = d->N, H = d->H, W = d->W, Cin = d->Cin, Cout = d->Cout, st_ = d->stride;
    CUtensorMap mx, mw;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        // traversal strides: with stride s the box spans TW*s x TH*s input pixels and TMA keeps every s-th -> TW x TH rows
        cuuint32_t box[4] = {kBlockK, (cuuint32_t)(kTileW * st_), (cuuint32_t)(kTileH * st_), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)st_, (cuuint32_t)st_, 1};
        CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail_status(MFB_ERR_CUDA, "conv: cuTensorMapEncodeTiled(x) failed with code " + std::to_string((int)r));
    }
    {
        const cuuint64_t ktot = (cuuint64_t)d->KH * d->KW * Cin;
        const int rank = d->per_image_weights ? 3 : 2;
        cuuint64_t dims[3] = {ktot, (cuuint64_t)Cout, (cuuint64_t)N};
        cuuint64_t strides[2] = {ktot * 2, ktot * 2 * (cuuint64_t)Cout};
        cuuint32_t box[3] = {kBlockK, (cuuint32_t)block_n, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(wgt), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail_status(MFB_ERR_CUDA, "conv: cuTensorMapEncodeTiled(w) failed with code " + std::to_string((int)r));
    }
    Params p;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Ho = d->Ho; p.Wo = d->Wo; p.Cout = Cout;
    p.KH = d->KH; p.KW = d->KW; p.stride = st_; p.pad_h = d->pad_h; p.pad_w = d->pad_w; p.act = d->act;
    p.tiles_w = (d->Wo + kTileW - 1) / kTileW;
    p.tiles_h = (d->Ho + kTileH - 1) / kTileH;
    p.per_image_w = d->per_image_weights ? 1 : 0;
    p.y = (__nv_bfloat16*)y; p.res = (const __nv_bfloat16*)residual;
    p.scale = (const float*)scale; p.shift = (const float*)shift;
    p.head_out = head ? (float*)head_out : nullptr;
    p.head_w = (const float*)head_w;
    for (int i = 0; i < kMaxHeads; ++i) {
        p.head_b[i] = d->head_bias[i]; p.head_lo[i] = d->head_lo[i]; p.head_hi[i] = d->head_hi[i]; p.head_act[i] = d->head_act[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (head) {
        if (d->act != kGelu) return fail_status(MFB_ERR_UNSUPPORTED, "conv: head mode is instantiated for the GELU heads of BevEncode only");
        return block_n == 128 ? launch<128, kGelu, true>(mx, mw, p, st) : launch<64, kGelu, true>(mx, mw, p, st);
    }
    if (wide_n) {
        switch (d->act) {
            case kNone: return launch256<kNone>(mx, mw, p, st);
            case kRelu: return launch256<kRelu>(mx, mw, p, st);
            case kGelu: return launch256<kGelu>(mx, mw, p, st);
            default: return launch256<kSilu>(mx, mw, p, st);
        }
    }
    switch (d->act * 2 + (block_n == 128)) {
        case kNone * 2: return launch<64, kNone, false>(mx, mw, p, st);
        case kNone * 2 + 1: return launch<128, kNone, false>(mx, mw, p, st);
        case kRelu * 2: return launch<64, kRelu, false>(mx, mw, p, st);
        case kRelu * 2 + 1: return launch<128, kRelu, false>(mx, mw, p, st);
        case kGelu * 2: return launch<64, kGelu, false>(mx, mw, p, st);
        case kGelu * 2 + 1: return launch<128, kGelu, false>(mx, mw, p, st);
        case kSilu * 2: return launch<64, kSilu, false>(mx, mw, p, st);
        default: return launch<128, kSilu, false>(mx, mw, p, st);
    }
}

// ABI v2 entry point: KS x KS, stride 1, "same" padding (kept for existing callers; forwards to mfb_conv2d_bf16)
extern "C" int mfb_conv_bn_act_bf16(const void* x, const void* wgt, const void* scale, const void* shift, void* y,
                                    int N, int H, int W, int Cin, int Cout, int KS, int act, void* stream) {
    if (KS != 1 && KS != 3) return fail_status(MFB_ERR_UNSUPPORTED, "conv: kernel size must be 1 or 3");
    if (act < kNone || act > kGelu) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: unknown activation");
    mfb_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.N = N; d.H = H; d.W = W; d.Cin = Cin; d.Ho = H; d.Wo = W; d.Cout = Cout;
    d.KH = d.KW = KS; d.stride = 1; d.pad_h = d.pad_w = KS / 2; d.act = act;
    return mfb_conv2d_bf16(&d, x, wgt, scale, shift, nullptr, y, nullptr, nullptr, stream);
}
