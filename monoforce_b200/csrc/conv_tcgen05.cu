// Host side of K4 (conv_tcgen05.cuh): TMA tensor maps + launch, exported through the C ABI.
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

template <int BLOCK_N>
static int launch(const CUtensorMap& mx, const CUtensorMap& mw, const Params& p, cudaStream_t st) {
    auto kern = conv_bn_act_kernel<BLOCK_N>;
    const int smem = Smem<BLOCK_N>::kTotal;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return fail_status(MFB_ERR_CUDA, "conv: cudaFuncSetAttribute failed");
    dim3 grid(p.tiles_w * p.tiles_h * p.N, p.Cout / BLOCK_N);
    kern<<<grid, kThreads, smem, st>>>(mx, mw, p);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_status(MFB_ERR_CUDA, std::string("conv launch: ") + cudaGetErrorString(e));
    return MFB_OK;
}

}  // namespace conv
}  // namespace mfb

using namespace mfb;
using namespace mfb::conv;

extern "C" int mfb_conv_bn_act_bf16(const void* x, const void* wgt, const void* scale, const void* shift, void* y,
                                    int N, int H, int W, int Cin, int Cout, int KS, int act, void* stream) {
    if (!x || !wgt || !scale || !shift || !y) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: NULL pointer");
    if (N < 1 || H < 1 || W < 1) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: sizes must be positive");
    if (KS != 1 && KS != 3) return fail_status(MFB_ERR_UNSUPPORTED, "conv: kernel size must be 1 or 3");
    if (Cin % kBlockK) return fail_status(MFB_ERR_UNSUPPORTED, "conv: Cin must be a multiple of 64 (pad the channels)");
    if (Cout % 64) return fail_status(MFB_ERR_UNSUPPORTED, "conv: Cout must be a multiple of 64");
    if (act < kNone || act > kGelu) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: unknown activation");
    if (((uintptr_t)x | (uintptr_t)wgt | (uintptr_t)y) & 15) return fail_status(MFB_ERR_INVALID_ARGUMENT, "conv: tensors must be 16-byte aligned");
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail_status(MFB_ERR_CUDA, "conv: cuTensorMapEncodeTiled is not available from the driver");

    const int block_n = (Cout % 128 == 0) ? 128 : 64;
    CUtensorMap mx, mw;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        cuuint32_t box[4] = {kBlockK, kTileW, kTileH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail_status(MFB_ERR_CUDA, "conv: cuTensorMapEncodeTiled(x) failed with code " + std::to_string((int)r));
    }
    {
        const cuuint64_t ktot = (cuuint64_t)KS * KS * Cin;
        cuuint64_t dims[2] = {ktot, (cuuint64_t)Cout};
        cuuint64_t strides[1] = {ktot * 2};
        cuuint32_t box[2] = {kBlockK, (cuuint32_t)block_n};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wgt), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail_status(MFB_ERR_CUDA, "conv: cuTensorMapEncodeTiled(w) failed with code " + std::to_string((int)r));
    }
    Params p;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.KS = KS; p.act = act;
    p.tiles_w = (W + kTileW - 1) / kTileW;
    p.tiles_h = (H + kTileH - 1) / kTileH;
    p.y = (__nv_bfloat16*)y; p.scale = (const float*)scale; p.shift = (const float*)shift;
    cudaStream_t st = (cudaStream_t)stream;
    return block_n == 128 ? launch<128>(mx, mw, p, st) : launch<64>(mx, mw, p, st);
}
