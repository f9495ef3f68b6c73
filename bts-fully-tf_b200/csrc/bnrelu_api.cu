// bnrelu_api.cu -- C ABI of the training-mode BatchNormalization (+ ReLU) glue over channel slices (include/btslpg.h:
// btslpg_bn_moments, btslpg_bn_fold, btslpg_bn_act_backward_stats, btslpg_bn_act_backward); one translation unit of libbtslpg.so.
#include "api_common.cuh"
#include "bnrelu_kernels.cuh"
#include "bnstat_kernels.cuh"   // kBnHeaderBytes / the shared workspace size

using namespace btslpg;
using namespace btslpg_api;

namespace {

// float32 channel slice whose 16-byte vectors are aligned; at most 1024 channels (a CTA of <= 256 threads covers whole pixels)
int parse_slice(const BtsTensor *t, const char *name, View &v) {
    if (int e = parse_pixel_strided(t, name, v)) return e;
    if (v.dtype != kF32) return fail(BTSLPG_EDTYPE, "%s: float32 only", name);
    if (v.C < 4 || v.C % 4 || v.C > 1024) return fail(BTSLPG_ESHAPE, "%s: %lld channels; a multiple of 4 in [4, 1024] is required", name, (long long)v.C);
    const int64_t npix = v.B * v.H * v.W;
    if (npix == 0) return fail(BTSLPG_ESHAPE, "%s: empty tensor", name);
    if (npix * v.C >= ((int64_t)1 << 33)) return fail(BTSLPG_ESHAPE, "%s: too many elements", name);
    if (!v.aligned(16) || (npix > 1 && v.sW % 4)) return fail(BTSLPG_ELAYOUT, "%s: pixels must start on 16-byte boundaries", name);
    return 0;
}
int64_t px_stride(const View &v) { return v.B * v.H * v.W == 1 ? v.C : v.sW; }

int same_pixels(const View &a, const View &b, const char *name) {
    if (a.B != b.B || a.H != b.H || a.W != b.W || a.C != b.C) return fail(BTSLPG_ESHAPE, "%s: shape differs", name);
    if (a.dev != b.dev) return fail(BTSLPG_EDEVICE, "%s: on a different device", name);
    return 0;
}

void reduce_launch_shape(uint64_t npix, uint32_t C, int &threads, int &blocks) {
    const uint32_t vpp = C / 4;
    threads = (int)((kBnrMaxThreads / vpp) * vpp);
    const uint32_t ppp = threads / vpp;
    uint64_t b = (npix + (uint64_t)ppp * 16 - 1) / ((uint64_t)ppp * 16);        // at least 16 passes per CTA
    if (b > (uint64_t)kBnrMaxBlocks) b = kBnrMaxBlocks;
    blocks = b < 1 ? 1 : (int)b;
}

int check_ws(void *workspace, size_t workspace_bytes, int64_t C, const char *fn) {
    if (!workspace || workspace_bytes < btslpg_bn_workspace_bytes((int)C) || (reinterpret_cast<uintptr_t>(workspace) % 16))
        return fail(BTSLPG_EWORKSPACE, "%s: workspace of btslpg_bn_workspace_bytes(C) bytes, 16-byte aligned, is required", fn);
    return 0;
}

struct Vecs {
    float *scale = nullptr, *shift = nullptr, *mean = nullptr, *rstd = nullptr;
};
int parse_vecs(const BtsTensor *scale, const BtsTensor *shift, const BtsTensor *mean, const BtsTensor *rstd, int64_t C, int dev, Vecs &v) {
    if (int e = parse_f32_vec(scale, "scale", C, dev, v.scale)) return e;
    if (int e = parse_f32_vec(shift, "shift", C, dev, v.shift)) return e;
    if (int e = parse_f32_vec(mean, "mean", C, dev, v.mean)) return e;
    if (int e = parse_f32_vec(rstd, "rstd", C, dev, v.rstd)) return e;
    for (const float *p : {v.scale, v.shift, v.mean, v.rstd})
        if (reinterpret_cast<uintptr_t>(p) % 16) return fail(BTSLPG_ELAYOUT, "per-channel vectors must be 16-byte aligned (slice them at multiples of 4 channels)");
    return 0;
}

}  // namespace

extern "C" {

int btslpg_bn_moments(const BtsTensor *x, BtsTensor *mean, BtsTensor *var, void *workspace, size_t workspace_bytes, void *stream) {
    View xv;
    if (int e = parse_slice(x, "x", xv)) return e;
    float *m = nullptr, *v = nullptr;
    if (int e = parse_f32_vec(mean, "mean", xv.C, xv.dev, m)) return e;
    if (int e = parse_f32_vec(var, "var", xv.C, xv.dev, v)) return e;
    if (int e = check_ws(workspace, workspace_bytes, xv.C, "btslpg_bn_moments")) return e;
    DeviceGuard guard(xv.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", xv.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BnrReduceParams p;
    memset(&p, 0, sizeof(p));
    p.x = reinterpret_cast<const float *>(xv.ptr); p.sx = px_stride(xv);
    p.npix = (uint64_t)(xv.B * xv.H * xv.W); p.C = (uint32_t)xv.C;
    p.partial = reinterpret_cast<double *>(static_cast<char *>(workspace) + kBnHeaderBytes);
    p.out0 = m; p.out1 = v;
    int threads, blocks;
    reduce_launch_shape(p.npix, p.C, threads, blocks);
    bnr_reduce_kernel<0><<<blocks, threads, threads * 8 * sizeof(double), st>>>(p);
    if (int e = check_launch("btslpg_bn_moments")) return e;
    bnr_finalize_kernel<0><<<(p.C + 7) / 8, 256, 0, st>>>(p, (uint32_t)blocks);
    snprintf(tl_kernel, sizeof(tl_kernel), "bn_moments<f32,C%u>", p.C);
    return check_launch("btslpg_bn_moments");
}

int btslpg_bn_fold(const BtsTensor *mean, const BtsTensor *var, const BtsTensor *gamma, const BtsTensor *beta, BtsTensor *running_mean,
                   BtsTensor *running_var, float momentum, float eps, int64_t count, BtsTensor *scale, BtsTensor *shift, BtsTensor *rstd,
                   void *stream) {
    View gv;
    if (int e = parse_common(gamma, "gamma", gv)) return e;
    int64_t C = 1;
    for (int k = 0; k < gamma->ndim; ++k) C *= gamma->shape[k];
    if (C < 1) return fail(BTSLPG_ESHAPE, "gamma: empty");
    if (count < 1) return fail(BTSLPG_EINVAL, "count must be the number of values per channel (>= 1)");
    BnrFoldParams p;
    memset(&p, 0, sizeof(p));
    float *m, *v, *g, *b, *sc, *sh, *rs;
    if (int e = parse_f32_vec(mean, "mean", C, gv.dev, m)) return e;
    if (int e = parse_f32_vec(var, "var", C, gv.dev, v)) return e;
    if (int e = parse_f32_vec(gamma, "gamma", C, gv.dev, g)) return e;
    if (int e = parse_f32_vec(beta, "beta", C, gv.dev, b)) return e;
    if (int e = parse_f32_vec(scale, "scale", C, gv.dev, sc)) return e;
    if (int e = parse_f32_vec(shift, "shift", C, gv.dev, sh)) return e;
    if (int e = parse_f32_vec(rstd, "rstd", C, gv.dev, rs)) return e;
    if (running_mean || running_var) {
        if (!running_mean || !running_var) return fail(BTSLPG_EINVAL, "running_mean and running_var go together");
        if (int e = parse_f32_vec(running_mean, "running_mean", C, gv.dev, p.running_mean)) return e;
        if (int e = parse_f32_vec(running_var, "running_var", C, gv.dev, p.running_var)) return e;
    }
    p.mean = m; p.var = v; p.gamma = g; p.beta = b; p.scale = sc; p.shift = sh; p.rstd = rs;
    p.momentum = momentum; p.eps = eps; p.count = (double)count; p.C = (uint32_t)C;
    DeviceGuard guard(gv.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", gv.dev, cudaGetErrorString(guard.err));
    bnr_fold_kernel<<<(p.C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    snprintf(tl_kernel, sizeof(tl_kernel), "bn_fold<C%u>", p.C);
    return check_launch("btslpg_bn_fold");
}

int btslpg_bn_act_backward_stats(const BtsTensor *g, const BtsTensor *g2, const BtsTensor *x, const BtsTensor *scale, const BtsTensor *shift,
                                 const BtsTensor *mean, const BtsTensor *rstd, int relu, BtsTensor *g_gamma, BtsTensor *g_beta,
                                 void *workspace, size_t workspace_bytes, void *stream) {
    View gv, hv, xv;
    if (int e = parse_slice(g, "g", gv)) return e;
    if (int e = parse_slice(x, "x", xv)) return e;
    if (int e = same_pixels(gv, xv, "x")) return e;
    if (g2) {
        if (int e = parse_slice(g2, "g2", hv)) return e;
        if (int e = same_pixels(gv, hv, "g2")) return e;
    }
    Vecs vc;
    if (int e = parse_vecs(scale, shift, mean, rstd, gv.C, gv.dev, vc)) return e;
    float *gg = nullptr, *gb = nullptr;
    if (int e = parse_f32_vec(g_gamma, "g_gamma", gv.C, gv.dev, gg)) return e;
    if (int e = parse_f32_vec(g_beta, "g_beta", gv.C, gv.dev, gb)) return e;
    if (int e = check_ws(workspace, workspace_bytes, gv.C, "btslpg_bn_act_backward_stats")) return e;
    DeviceGuard guard(gv.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", gv.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BnrReduceParams p;
    memset(&p, 0, sizeof(p));
    p.x = reinterpret_cast<const float *>(xv.ptr); p.sx = px_stride(xv);
    p.g = reinterpret_cast<const float *>(gv.ptr); p.sg = px_stride(gv);
    if (g2) { p.g2 = reinterpret_cast<const float *>(hv.ptr); p.sg2 = px_stride(hv); }
    p.scale = vc.scale; p.shift = vc.shift; p.mean = vc.mean; p.rstd = vc.rstd;
    p.relu = relu ? 1 : 0;
    p.npix = (uint64_t)(gv.B * gv.H * gv.W); p.C = (uint32_t)gv.C;
    p.partial = reinterpret_cast<double *>(static_cast<char *>(workspace) + kBnHeaderBytes);
    p.out0 = gb; p.out1 = gg;
    int threads, blocks;
    reduce_launch_shape(p.npix, p.C, threads, blocks);
    bnr_reduce_kernel<1><<<blocks, threads, threads * 8 * sizeof(double), st>>>(p);
    if (int e = check_launch("btslpg_bn_act_backward_stats")) return e;
    bnr_finalize_kernel<1><<<(p.C + 7) / 8, 256, 0, st>>>(p, (uint32_t)blocks);
    snprintf(tl_kernel, sizeof(tl_kernel), "bn_act_bwd_stats<f32,C%u,%s>", p.C, relu ? "relu" : "id");
    return check_launch("btslpg_bn_act_backward_stats");
}

int btslpg_bn_act_backward(const BtsTensor *g, const BtsTensor *g2, const BtsTensor *x, const BtsTensor *scale, const BtsTensor *shift,
                           const BtsTensor *mean, const BtsTensor *rstd, const BtsTensor *g_gamma, const BtsTensor *g_beta, int relu,
                           BtsTensor *dst, int accumulate, const BtsTensor *dst_init, void *stream) {
    View gv, hv, xv, dv, iv;
    if (dst_init) {
        if (accumulate) return fail(BTSLPG_EINVAL, "dst_init (dst = dst_init + value) and accumulate (dst += value) exclude each other");
        if (int e = parse_slice(dst_init, "dst_init", iv)) return e;
    }
    if (int e = parse_slice(g, "g", gv)) return e;
    if (int e = parse_slice(x, "x", xv)) return e;
    if (int e = parse_slice(dst, "dst", dv)) return e;
    if (int e = same_pixels(gv, xv, "x")) return e;
    if (int e = same_pixels(gv, dv, "dst")) return e;
    if (dst_init) {
        if (int e = same_pixels(gv, iv, "dst_init")) return e;
    }
    if (g2) {
        if (int e = parse_slice(g2, "g2", hv)) return e;
        if (int e = same_pixels(gv, hv, "g2")) return e;
    }
    Vecs vc;
    if (int e = parse_vecs(scale, shift, mean, rstd, gv.C, gv.dev, vc)) return e;
    float *gg = nullptr, *gb = nullptr;
    if (int e = parse_f32_vec(g_gamma, "g_gamma", gv.C, gv.dev, gg)) return e;
    if (int e = parse_f32_vec(g_beta, "g_beta", gv.C, gv.dev, gb)) return e;
    if ((reinterpret_cast<uintptr_t>(gg) % 16) || (reinterpret_cast<uintptr_t>(gb) % 16))
        return fail(BTSLPG_ELAYOUT, "g_gamma / g_beta must be 16-byte aligned");
    const int64_t npix = gv.B * gv.H * gv.W;
    if (npix * (gv.C / 4) >= ((int64_t)1 << 31)) return fail(BTSLPG_ESHAPE, "g: more than 2^31 16-byte vectors");
    DeviceGuard guard(gv.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", gv.dev, cudaGetErrorString(guard.err));
    BnrApplyParams p;
    memset(&p, 0, sizeof(p));
    p.g = reinterpret_cast<const float *>(gv.ptr); p.sg = px_stride(gv);
    if (g2) { p.g2 = reinterpret_cast<const float *>(hv.ptr); p.sg2 = px_stride(hv); }
    p.x = reinterpret_cast<const float *>(xv.ptr); p.sx = px_stride(xv);
    p.dst = reinterpret_cast<float *>(dv.ptr); p.sd = px_stride(dv);
    if (dst_init) { p.init = reinterpret_cast<const float *>(iv.ptr); p.si = px_stride(iv); }
    p.scale = vc.scale; p.shift = vc.shift; p.mean = vc.mean; p.rstd = vc.rstd; p.g_beta = gb; p.g_gamma = gg;
    p.inv_n = (float)(1.0 / (double)npix);
    p.relu = relu ? 1 : 0; p.accumulate = accumulate ? 1 : 0;
    p.vpp = (uint32_t)(gv.C / 4);
    p.n = (uint64_t)npix * p.vpp;
    p.div_vpp = FastDiv(p.vpp);
    bnr_apply_kernel<<<(unsigned)((p.n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    snprintf(tl_kernel, sizeof(tl_kernel), "bn_act_bwd<f32,C%u,%s%s>", (unsigned)gv.C, relu ? "relu" : "id", accumulate ? ",acc" : "");
    return check_launch("btslpg_bn_act_backward");
}

}  // extern "C"
