// bnstat_api.cu -- C ABI of the training-mode conv-block glue statistics (include/btslpg.h: btslpg_bn_elu_stats,
// btslpg_bn_elu_backward_stats); one translation unit of libbtslpg.so.
#include "api_common.cuh"
#include "bnstat_kernels.cuh"

using namespace btslpg;
using namespace btslpg_api;

namespace {

bool bn_channels_ok(int64_t C) { return C >= 4 && C <= 1024 && (C & (C - 1)) == 0; }

int bn_blocks(uint64_t npix, uint32_t C) {
    const uint32_t px_per_pass = kBnThreads / (C / 4);
    uint64_t b = (npix + (uint64_t)px_per_pass * 16 - 1) / ((uint64_t)px_per_pass * 16);    // at least 16 passes per CTA
    if (b > (uint64_t)kBnMaxBlocks) b = kBnMaxBlocks;
    return b < 1 ? 1 : (int)b;
}

}  // namespace

extern "C" {

size_t btslpg_bn_workspace_bytes(int channels) {
    if (channels < 1) channels = 1;
    return (size_t)kBnHeaderBytes + (size_t)kBnMaxBlocks * 2 * channels * sizeof(double);
}

int btslpg_bn_elu_stats(const BtsTensor *raw, int act, const BtsTensor *gamma, const BtsTensor *beta, BtsTensor *running_mean,
                        BtsTensor *running_var, float momentum, float eps, BtsTensor *pack, void *workspace, size_t workspace_bytes,
                        void *stream) {
    View rv;
    if (int e = parse_nhwc(raw, "raw", rv)) return e;
    if (rv.dtype != kF32) return fail(BTSLPG_EDTYPE, "raw: float32 only");
    if (!is_contig_nhwc(rv) || !rv.aligned(16)) return fail(BTSLPG_ELAYOUT, "raw: must be contiguous NHWC and 16-byte aligned");
    const int64_t C = rv.C;
    if (!bn_channels_ok(C)) return fail(BTSLPG_ESHAPE, "raw: %lld channels; the statistics kernel takes powers of two in [4, 1024]", (long long)C);
    float *g = nullptr, *b = nullptr, *rm = nullptr, *rvar = nullptr, *pk = nullptr;
    if (int e = parse_f32_vec(gamma, "gamma", C, rv.dev, g)) return e;
    if (int e = parse_f32_vec(beta, "beta", C, rv.dev, b)) return e;
    if (running_mean || running_var) {
        if (!running_mean || !running_var) return fail(BTSLPG_EINVAL, "running_mean and running_var go together");
        if (int e = parse_f32_vec(running_mean, "running_mean", C, rv.dev, rm)) return e;
        if (int e = parse_f32_vec(running_var, "running_var", C, rv.dev, rvar)) return e;
    }
    if (int e = parse_f32_vec(pack, "pack", 8 * C, rv.dev, pk)) return e;
    const uint64_t npix = (uint64_t)(rv.B * rv.H * rv.W);
    if (npix == 0) return fail(BTSLPG_ESHAPE, "raw: batch statistics of an empty tensor");
    if (!workspace || workspace_bytes < btslpg_bn_workspace_bytes((int)C) || (reinterpret_cast<uintptr_t>(workspace) % 16))
        return fail(BTSLPG_EWORKSPACE, "bn_elu_stats: workspace of btslpg_bn_workspace_bytes(C) bytes, 16-byte aligned, is required");
    DeviceGuard guard(rv.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", rv.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BnStatParams p;
    memset(&p, 0, sizeof(p));
    p.a = reinterpret_cast<const float *>(rv.ptr);
    p.npix = npix; p.C = (uint32_t)C; p.stride = (uint32_t)C; p.act = act ? 1 : 0;
    p.gamma = g; p.beta = b; p.running_mean = rm; p.running_var = rvar; p.momentum = momentum; p.eps = eps; p.pack = pk;
    p.partial = reinterpret_cast<double *>(static_cast<char *>(workspace) + kBnHeaderBytes);
    const int blocks = bn_blocks(npix, p.C);
    bn_stats_kernel<false><<<blocks, kBnThreads, kBnThreads * 8 * sizeof(double), st>>>(p);
    if (int e = check_launch("btslpg_bn_elu_stats")) return e;
    bn_finalize_kernel<false><<<(p.C + 7) / 8, 256, 0, st>>>(p, (uint32_t)blocks);
    snprintf(tl_kernel, sizeof(tl_kernel), "bn_elu_stats<f32,C%u>", p.C);
    return check_launch("btslpg_bn_elu_stats");
}

int btslpg_bn_elu_backward_stats(const BtsTensor *g_out, const BtsTensor *y, int channels, BtsTensor *pack, BtsTensor *g_gamma, BtsTensor *g_beta,
                                 void *workspace, size_t workspace_bytes, void *stream) {
    View gv, yv;
    if (int e = parse_nhwc(g_out, "g_out", gv)) return e;
    if (int e = parse_nhwc(y, "y", yv)) return e;
    if (gv.dtype != kF32 || yv.dtype != kF32) return fail(BTSLPG_EDTYPE, "g_out / y: float32 only");
    if (!is_contig_nhwc(gv) || !is_contig_nhwc(yv) || !gv.aligned(16) || !yv.aligned(16))
        return fail(BTSLPG_ELAYOUT, "g_out / y: must be contiguous NHWC and 16-byte aligned");
    if (yv.B != gv.B || yv.H != gv.H || yv.W != gv.W || yv.C != gv.C) return fail(BTSLPG_ESHAPE, "y: shape differs from g_out");
    if (yv.dev != gv.dev) return fail(BTSLPG_EDEVICE, "y: on a different device than g_out");
    const int64_t C = channels;
    if (!bn_channels_ok(C) || C > gv.C) return fail(BTSLPG_ESHAPE, "channels = %lld: a power of two in [4, 1024], at most the concat width", (long long)C);
    if (gv.C % 4) return fail(BTSLPG_ELAYOUT, "g_out: the concat width must be a multiple of 4 (pad channels)");
    float *pk = nullptr, *gg = nullptr, *gb = nullptr;
    if (int e = parse_f32_vec(pack, "pack", 8 * C, gv.dev, pk)) return e;
    if (int e = parse_f32_vec(g_gamma, "g_gamma", C, gv.dev, gg)) return e;
    if (int e = parse_f32_vec(g_beta, "g_beta", C, gv.dev, gb)) return e;
    const uint64_t npix = (uint64_t)(gv.B * gv.H * gv.W);
    if (npix == 0) return fail(BTSLPG_ESHAPE, "g_out: empty tensor");
    if (!workspace || workspace_bytes < btslpg_bn_workspace_bytes((int)C) || (reinterpret_cast<uintptr_t>(workspace) % 16))
        return fail(BTSLPG_EWORKSPACE, "bn_elu_backward_stats: workspace of btslpg_bn_workspace_bytes(C) bytes, 16-byte aligned, is required");
    DeviceGuard guard(gv.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", gv.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BnStatParams p;
    memset(&p, 0, sizeof(p));
    p.a = reinterpret_cast<const float *>(gv.ptr);
    p.y = reinterpret_cast<const float *>(yv.ptr);
    p.npix = npix; p.C = (uint32_t)C; p.stride = (uint32_t)gv.C;
    p.pack = pk; p.g_gamma = gg; p.g_beta = gb;
    p.partial = reinterpret_cast<double *>(static_cast<char *>(workspace) + kBnHeaderBytes);
    const int blocks = bn_blocks(npix, p.C);
    bn_stats_kernel<true><<<blocks, kBnThreads, kBnThreads * 8 * sizeof(double), st>>>(p);
    if (int e = check_launch("btslpg_bn_elu_backward_stats")) return e;
    bn_finalize_kernel<true><<<(p.C + 7) / 8, 256, 0, st>>>(p, (uint32_t)blocks);
    snprintf(tl_kernel, sizeof(tl_kernel), "bn_elu_bwd_stats<f32,C%u>", p.C);
    return check_launch("btslpg_bn_elu_backward_stats");
}

}  // extern "C"
