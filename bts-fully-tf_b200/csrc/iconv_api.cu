// iconv_api.cu -- C ABI of the tcgen05 implicit-GEMM iconv1 (include/btslpg.h: btslpg_iconv1_forward); one translation unit of libbtslpg.so.
#include "api_common.cuh"
#include "iconv_kernels.cuh"

using namespace btslpg;
using namespace btslpg_api;

namespace {
std::atomic<unsigned long long *> g_iconv_prof{nullptr};
}

extern "C" {

// tools only (not in include/btslpg.h): device buffer of >= 16 uint64 that subsequent launches add their roles' wait cycles to
__attribute__((visibility("default"))) void btslpg_debug_iconv1_profile(void *device_u64) {
    g_iconv_prof.store(static_cast<unsigned long long *>(device_u64));
}

int btslpg_iconv1_forward(const BtsTensor *a, int a_subpixel, const BtsTensor *const *planes, const BtsTensor *kernel, int act_out,
                          BtsTensor *out, void *stream) {
    View av, ov, pv[3];
    if (int e = parse_nhwc(a, "a", av)) return e;
    if (int e = parse_nhwc(out, "out", ov)) return e;
    if (av.dtype != kF32 || ov.dtype != kF32) return fail(BTSLPG_EDTYPE, "iconv1_forward: float32 tensors only (TF32 tensor-core arithmetic)");
    const int64_t NF = ov.C;
    if (NF != 16 && NF != 32) return fail(BTSLPG_ESHAPE, "out: F/16 = %lld filters; the fused kernel has 16 and 32", (long long)NF);
    if (!is_contig_nhwc(av) || !is_contig_nhwc(ov) || !av.aligned(32) || !ov.aligned(32))
        return fail(BTSLPG_ELAYOUT, "a / out: must be contiguous NHWC and 32-byte aligned");
    if (a_subpixel) {
        if (av.B != ov.B || av.H * 2 != ov.H || av.W * 2 != ov.W || av.C != 4 * NF)
            return fail(BTSLPG_ESHAPE, "a: expected (B,H/2,W/2,4*%lld) for a_subpixel", (long long)NF);
    } else if (av.B != ov.B || av.H != ov.H || av.W != ov.W || av.C != NF) {
        return fail(BTSLPG_ESHAPE, "a: expected (B,H,W,%lld) like out", (long long)NF);
    }
    if (av.dev != ov.dev) return fail(BTSLPG_EDEVICE, "a: on a different device than out");
    if (!planes) return fail(BTSLPG_EINVAL, "planes: NULL");
    for (int k = 0; k < 3; ++k) {
        int64_t n = 0;
        char nm[16];
        snprintf(nm, sizeof(nm), "planes[%d]", k);
        if (int e = parse_flat(planes[k], nm, pv[k], n)) return e;
        if (pv[k].dtype != kF32) return fail(BTSLPG_EDTYPE, "%s: must be float32", nm);
        if (pv[k].B != ov.B || pv[k].H != ov.H || pv[k].W != ov.W) return fail(BTSLPG_ESHAPE, "%s: expected (B,H,W[,1]) like out", nm);
        if (pv[k].dev != ov.dev) return fail(BTSLPG_EDEVICE, "%s: on a different device than out", nm);
    }
    float *w = nullptr;
    if (int e = parse_f32_vec(kernel, "kernel", 9 * (NF + 3) * NF, ov.dev, w)) return e;
    if (ov.B * ov.H * ov.W == 0) return 0;
    if (ov.H >= (1 << 24) || ov.W >= (1 << 24)) return fail(BTSLPG_ESHAPE, "out: extent too large");

    IconvParams p;
    p.a = reinterpret_cast<const float *>(av.ptr);
    p.p0 = reinterpret_cast<const float *>(pv[0].ptr);
    p.p1 = reinterpret_cast<const float *>(pv[1].ptr);
    p.p2 = reinterpret_cast<const float *>(pv[2].ptr);
    p.w = w;
    p.out = reinterpret_cast<float *>(ov.ptr);
    p.B = (int)ov.B; p.H = (int)ov.H; p.W = (int)ov.W;
    p.a_subpixel = a_subpixel ? 1 : 0;
    p.act_out = act_out ? 1 : 0;
    p.nstrips = (p.W + kIcTW - 1) / kIcTW;
    p.strip_w = (p.W + p.nstrips - 1) / p.nstrips;
    p.rows_per_item = p.H < 32 ? p.H : 32;
    p.nseg = (p.H + p.rows_per_item - 1) / p.rows_per_item;
    const int64_t items = (int64_t)p.B * p.nseg * p.nstrips;
    if (items >= ((int64_t)1 << 31)) return fail(BTSLPG_ESHAPE, "out: too many work items");
    p.items = (uint32_t)items;
    p.prof = g_iconv_prof.load();

    DeviceGuard guard(ov.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", ov.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto go = [&](auto tag) -> int {
        constexpr int NFc = decltype(tag)::value;
        constexpr int smem = IconvCfg<NFc>::kSmemBytes;
        static PerDevice per_dev;
        const int resident = per_dev.get([&] { return occupancy_blocks_smem(iconv1_fwd_kernel<NFc>, kIcThreads, smem); });
        const uint32_t blocks = p.items < (uint32_t)resident ? p.items : (uint32_t)resident;
        iconv1_fwd_kernel<NFc><<<blocks, kIcThreads, smem, st>>>(p);
        snprintf(tl_kernel, sizeof(tl_kernel), "iconv1_fwd_tcgen05<f32/tf32,NF%d,%s%s>", NFc, a_subpixel ? "subpixel" : "nhwc", act_out ? ",elu" : "");
        return check_launch("btslpg_iconv1_forward");
    };
    return NF == 32 ? go(IntC<32>{}) : go(IntC<16>{});
}

}  // extern "C"
