// wgrad_api.cu -- C ABI of the tcgen05 weight gradient of the full-resolution 3x3 convolutions (include/btslpg.h:
// btslpg_conv3x3_wgrad); one translation unit of libbtslpg.so.
#include "api_common.cuh"
#include "wgrad_kernels.cuh"

using namespace btslpg;
using namespace btslpg_api;

namespace {

constexpr int kWgHeaderBytes = 256;   // the workspace's first 256 bytes belong to the kernels that keep counters there (always left zero)
constexpr int kWgMaxBlocks = 160;      // partial rows the workspace holds (>= SMs of the device; the launch uses min(items, SMs))

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// (B,H,W,C) float32 NHWC tensor as a 4-D map (C, W, H, B) with boxes of 32 channels x box_w pixels x box_h rows in the 128-byte-span / 32-byte-atom
// swizzle; elements outside the tensor (image borders, channels beyond C) read as zero
int make_map(const View &v, const char *name, int box_w, int box_h, CUtensorMap &map) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return fail(BTSLPG_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.B};
    const cuuint64_t strides[3] = {(cuuint64_t)v.C * 4, (cuuint64_t)v.W * v.C * 4, (cuuint64_t)v.H * v.W * v.C * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, v.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(BTSLPG_ECUDA, "%s: cuTensorMapEncodeTiled failed (CUresult %d)", name, (int)r);
    return 0;
}

}  // namespace

extern "C" {

size_t btslpg_conv3x3_wgrad_workspace_bytes(int cin, int cout) {
    if (cin < 1) cin = 1;
    if (cout < 1) cout = 1;
    return (size_t)kWgHeaderBytes + (size_t)kWgMaxBlocks * 9 * cin * cout * sizeof(float);
}

int btslpg_conv3x3_wgrad(const BtsTensor *x, const BtsTensor *g, BtsTensor *g_kernel, void *workspace, size_t workspace_bytes, void *stream) {
    View xv, gv;
    if (int e = parse_nhwc(x, "x", xv)) return e;
    if (int e = parse_nhwc(g, "g", gv)) return e;
    if (xv.dtype != kF32 || gv.dtype != kF32) return fail(BTSLPG_EDTYPE, "conv3x3_wgrad: float32 tensors only (TF32 tensor-core arithmetic)");
    if (!is_contig_nhwc(xv) || !is_contig_nhwc(gv) || !xv.aligned(16) || !gv.aligned(16))
        return fail(BTSLPG_ELAYOUT, "x / g: must be contiguous NHWC and 16-byte aligned");
    if (xv.B != gv.B || xv.H != gv.H || xv.W != gv.W) return fail(BTSLPG_ESHAPE, "g: expected (B,H,W,Cout) over the same pixels as x");
    if (xv.dev != gv.dev) return fail(BTSLPG_EDEVICE, "g: on a different device than x");
    const int64_t Cin = xv.C, Cout = gv.C;
    if (Cin < 4 || Cin > 256 || Cin % 4) return fail(BTSLPG_ESHAPE, "x: %lld channels; a multiple of 4 in [4, 256] is required", (long long)Cin);
    if (Cout < 4 || Cout > 128 || Cout % 4) return fail(BTSLPG_ESHAPE, "g: %lld channels; a multiple of 4 in [4, 128] is required", (long long)Cout);
    float *out = nullptr;
    if (int e = parse_f32_vec(g_kernel, "g_kernel", 9 * Cin * Cout, xv.dev, out)) return e;
    if (xv.B * xv.H * xv.W == 0) return fail(BTSLPG_ESHAPE, "x: empty tensor");
    if (xv.H >= (1 << 24) || xv.W >= (1 << 24) || xv.B >= (1 << 24)) return fail(BTSLPG_ESHAPE, "x: extent too large");
    if (!workspace || workspace_bytes < btslpg_conv3x3_wgrad_workspace_bytes((int)Cin, (int)Cout) || (reinterpret_cast<uintptr_t>(workspace) % 16))
        return fail(BTSLPG_EWORKSPACE, "conv3x3_wgrad: workspace of btslpg_conv3x3_wgrad_workspace_bytes(Cin, Cout) bytes, 16-byte aligned, is required");

    DeviceGuard guard(xv.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", xv.dev, cudaGetErrorString(guard.err));
    const int T = (Cin <= 32 && Cout <= 32) ? 16 : 8;                     // WgradCfg::kT of the pass shapes used below
    CUtensorMap map_x, map_g;
    if (int e = make_map(xv, "x", kWgTW, T + 2, map_x)) return e;
    if (int e = make_map(gv, "g", kWgTW + 2, T, map_g)) return e;

    WgradParams p;
    p.partial = reinterpret_cast<float *>(static_cast<char *>(workspace) + kWgHeaderBytes);
    p.B = (int)xv.B; p.H = (int)xv.H; p.W = (int)xv.W; p.Cin = (int)Cin; p.Cout = (int)Cout;
    p.bands = (uint32_t)((xv.H + T - 1) / T);
    p.ctiles = (uint32_t)((xv.W + kWgTW - 1) / kWgTW);
    const int64_t items = (int64_t)xv.B * p.bands * p.ctiles;
    if (items >= ((int64_t)1 << 31)) return fail(BTSLPG_ESHAPE, "x: too many work items");
    p.items = (uint32_t)items;
    p.div_ct = FastDiv(p.ctiles);
    p.div_band = FastDiv(p.bands);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static PerDevice per_dev_sms;
    const int sms = per_dev_sms.get([&] {
        int n = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n < 1 ? 1 : n;
    });
    uint32_t blocks = p.items < (uint32_t)sms ? p.items : (uint32_t)sms;
    if (blocks > (uint32_t)kWgMaxBlocks) blocks = kWgMaxBlocks;
    auto go = [&](auto cinb_tag, auto coutb_tag, auto t_tag) -> int {
        constexpr int CINB = decltype(cinb_tag)::value, COUTB = decltype(coutb_tag)::value, TT = decltype(t_tag)::value;
        constexpr int smem = WgradCfg<CINB, COUTB, TT>::kSmemBytes;
        static PerDevice per_dev;
        per_dev.get([&] {
            cudaFuncSetAttribute(conv3x3_wgrad_kernel<CINB, COUTB, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            return 1;
        });
        conv3x3_wgrad_kernel<CINB, COUTB, TT><<<blocks, kWgThreads, smem, st>>>(map_x, map_g, p);
        return check_launch("btslpg_conv3x3_wgrad");
    };
    // channels in passes of up to 64 x 64 (two 32-channel blocks each: the accumulators of a pass must fit the 512 TMEM columns; every
    // pass stages its own operands, so an operand with more than 64 channels makes the other one be read again)
    const bool small = Cin <= 32 && Cout <= 32;
    for (int ci0 = 0; ci0 < (int)Cin; ci0 += 64)
        for (int co0 = 0; co0 < (int)Cout; co0 += 64) {
            p.ci0 = ci0; p.ci_n = (int)Cin - ci0 < 64 ? (int)Cin - ci0 : 64;
            p.co0 = co0; p.co_n = (int)Cout - co0 < 64 ? (int)Cout - co0 : 64;
            // one shape per call (the tensor maps' boxes are per call): 1 x 1 blocks only when both operands are narrow
            const bool two_in = !small && p.ci_n > 32, two_out = !small && p.co_n > 32;
            int e;
            if (small) e = go(IntC<1>{}, IntC<1>{}, IntC<16>{});
            else if (two_in) e = two_out ? go(IntC<2>{}, IntC<2>{}, IntC<8>{}) : go(IntC<2>{}, IntC<1>{}, IntC<8>{});
            else e = two_out ? go(IntC<1>{}, IntC<2>{}, IntC<8>{}) : go(IntC<1>{}, IntC<1>{}, IntC<8>{});
            if (e) return e;
        }
    const uint32_t n = (uint32_t)(9 * Cin * Cout);
    wgrad_reduce_kernel<<<(n + 31) / 32, 256, 0, st>>>(p.partial, out, n, blocks);
    snprintf(tl_kernel, sizeof(tl_kernel), "conv3x3_wgrad_tcgen05<f32/tf32,Cin%d,Cout%d>", p.Cin, p.Cout);
    return check_launch("btslpg_conv3x3_wgrad");
}

}  // extern "C"
