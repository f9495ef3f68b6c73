// bnrelu_kernels.cuh -- TRAINING-mode BatchNormalization (+ ReLU) over channel slices of NHWC buffers (sm_100a).
//
// bts_decoder.py:46-54 (dense_aspp_block) and :61-76: every block of the DenseASPP runs
//     [Concatenate ->] BatchNormalization(training) -> ReLU -> Conv2D 1x1 -> BatchNormalization(training) -> ReLU -> dilated Conv2D
// on a feature map that grows by Concatenate from 256 to 896 channels.  The framework copies the growing map for every concat, reads
// it once for the batch statistics, reads and writes it for the normalisation and again for the ReLU -- and runs the mirror image of
// all that backward.  Two facts remove most of it:
//   * batch statistics are PER CHANNEL, and the blocks' inputs share channels (concat4_k = first Ck channels of one buffer): the mean
//     and variance of a channel are computed once, when the channel is appended, and reused by every later BatchNormalization
//     (which differ only in gamma / beta);
//   * with the folded affine (scale = gamma * rstd, shift = beta - mean * scale) the forward is the inference pass of
//     slice_kernels.cuh: ONE read of the slice, one write of the 1x1 conv's contiguous input.
// Backward of y = relu(x * scale + shift), with gm = g * [y > 0] (+ g2, a gradient that reaches the normalised value directly):
//     d beta = sum gm ; d gamma = sum gm * xhat ; d x = scale * (gm - d beta / n - xhat * d gamma / n),  xhat = (x - mean) * rstd
// is a reduction pass (reads g and x) and an apply pass (reads g and x, adds into the gradient slice of the shared buffer: the concat's
// backward is that accumulation, no slicing copies).  The ReLU mask is recomputed with the forward's own fmaf.
//
// Reductions are deterministic: a thread owns four channels, fixed-order sums thread -> CTA (float64 in shared memory) -> grid (rows of
// float64 partials summed in a fixed order by the finalize kernel); no atomics.
#pragma once

#include "common.cuh"

namespace btslpg {

constexpr int kBnrMaxThreads = 256;
constexpr int kBnrMaxBlocks = 148 * 4;

struct BnrReduceParams {
    const float *x;   int64_t sx;    // MODE 0: the values; MODE 1: the BatchNormalization's input.  Pixel strides in elements.
    const float *g;   int64_t sg;    // MODE 1: gradient of the activation's output
    const float *g2;  int64_t sg2;   // MODE 1, nullable: gradient reaching the normalised value directly (not masked)
    const float *scale, *shift, *mean, *rstd;
    int relu;
    uint64_t npix;
    uint32_t C;
    double *partial;                 // [gridDim.x][2][C]
    float *out0, *out1;              // MODE 0: mean, biased variance; MODE 1: d beta, d gamma
};

// MODE 0: sums of x and x^2;  MODE 1: sums of gm and gm * xhat
template <int MODE> __global__ void __launch_bounds__(kBnrMaxThreads) bnr_reduce_kernel(const __grid_constant__ BnrReduceParams prm) {
    extern __shared__ __align__(16) unsigned char bnr_smem[];
    double *red = reinterpret_cast<double *>(bnr_smem);                 // [blockDim.x][8]
    const uint32_t C = prm.C, vpp = C / 4, nthr = blockDim.x;           // nthr is a multiple of vpp (host)
    const uint32_t ppp = nthr / vpp;                                     // pixels per pass
    const uint32_t cbase = 4 * (threadIdx.x % vpp);
    float sc[4], sh[4], mu[4], rs[4];
    if (MODE == 1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sc[e] = __ldg(prm.scale + cbase + e); sh[e] = __ldg(prm.shift + cbase + e);
            mu[e] = __ldg(prm.mean + cbase + e);  rs[e] = __ldg(prm.rstd + cbase + e);
        }
    }
    const uint64_t px0 = (uint64_t)blockIdx.x * ppp + threadIdx.x / vpp;
    const uint64_t step = (uint64_t)gridDim.x * ppp;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    double d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
    uint32_t run = 0;
    constexpr int U = 4;                                                 // pixels per thread and trip: all loads first
    for (uint64_t p = px0; p < prm.npix; p += U * step) {
        float4 xv[U], gv[U], hv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t q = p + u * step;
            if (q < prm.npix) {
                xv[u] = __ldg(reinterpret_cast<const float4 *>(prm.x + q * prm.sx + cbase));
                if (MODE == 1) {
                    gv[u] = __ldg(reinterpret_cast<const float4 *>(prm.g + q * prm.sg + cbase));
                    if (prm.g2) hv[u] = __ldg(reinterpret_cast<const float4 *>(prm.g2 + q * prm.sg2 + cbase));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (p + u * step >= prm.npix) break;
            const float x[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
            if (MODE == 0) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { s1[e] += x[e]; s2[e] = fmaf(x[e], x[e], s2[e]); }
            } else {
                const float g[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
                float h[4] = {0.f, 0.f, 0.f, 0.f};
                if (prm.g2) { h[0] = hv[u].x; h[1] = hv[u].y; h[2] = hv[u].z; h[3] = hv[u].w; }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool on = !prm.relu || fmaf(x[e], sc[e], sh[e]) > 0.0f;
                    const float gm = (on ? g[e] : 0.0f) + h[e];
                    s1[e] += gm;
                    s2[e] = fmaf(gm, (x[e] - mu[e]) * rs[e], s2[e]);
                }
            }
        }
        if (++run == 16) {                                               // float32 runs of 64 terms, float64 above: fixed order either way
#pragma unroll
            for (int e = 0; e < 4; ++e) { d1[e] += (double)s1[e]; d2[e] += (double)s2[e]; s1[e] = s2[e] = 0.f; }
            run = 0;
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        red[threadIdx.x * 8 + e] = d1[e] + (double)s1[e];
        red[threadIdx.x * 8 + 4 + e] = d2[e] + (double)s2[e];
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 2 * C; i += nthr) {               // channel 4 cg + e is owned by threads cg, cg + vpp, ...
        const uint32_t which = i / C, c = i % C, cg = c / 4, e = c % 4;
        double s = 0.0;
        for (uint32_t t = cg; t < nthr; t += vpp) s += red[t * 8 + which * 4 + e];
        prm.partial[(size_t)blockIdx.x * 2 * C + i] = s;
    }
}

// CTA rows -> result: 8 channels per CTA, 32 row slices per channel, slices combined in slice order (fixed order, no atomics)
template <int MODE> __global__ void __launch_bounds__(256) bnr_finalize_kernel(const __grid_constant__ BnrReduceParams prm, uint32_t nrows) {
    __shared__ double comb[32][8][2];
    const uint32_t C = prm.C, cl = threadIdx.x % 8, sl = threadIdx.x / 8, c = blockIdx.x * 8 + cl;
    double a1 = 0.0, a2 = 0.0;
    if (c < C) {
        uint32_t b = sl;
        for (; b + 96 < nrows; b += 128) {
            double u[4], w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u[k] = __ldcg(prm.partial + (size_t)(b + 32 * k) * 2 * C + c);
                w[k] = __ldcg(prm.partial + (size_t)(b + 32 * k) * 2 * C + C + c);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { a1 += u[k]; a2 += w[k]; }
        }
        for (; b < nrows; b += 32) {
            a1 += __ldcg(prm.partial + (size_t)b * 2 * C + c);
            a2 += __ldcg(prm.partial + (size_t)b * 2 * C + C + c);
        }
    }
    comb[sl][cl][0] = a1;
    comb[sl][cl][1] = a2;
    __syncthreads();
    if (sl != 0 || c >= C) return;
#pragma unroll
    for (int q = 1; q < 32; ++q) { a1 += comb[q][cl][0]; a2 += comb[q][cl][1]; }
    if (MODE == 0) {
        const double n = (double)prm.npix, mean = a1 / n;
        double var = a2 / n - mean * mean;                               // biased: what training-mode normalisation uses
        prm.out0[c] = (float)mean;
        prm.out1[c] = (float)(var < 0.0 ? 0.0 : var);
    } else {
        prm.out0[c] = (float)a1;
        prm.out1[c] = (float)a2;
    }
}

struct BnrFoldParams {
    const float *mean, *var, *gamma, *beta;
    float *running_mean, *running_var;     // nullable
    float momentum, eps;                   // torch convention: running = (1 - momentum) * running + momentum * batch
    double count;
    float *scale, *shift, *rstd;
    uint32_t C;
};

__global__ void __launch_bounds__(256) bnr_fold_kernel(const __grid_constant__ BnrFoldParams prm) {
    const uint32_t c = blockIdx.x * 256 + threadIdx.x;
    if (c >= prm.C) return;
    const double mean = prm.mean[c], var = prm.var[c];
    const double rstd = 1.0 / sqrt(var + (double)prm.eps);
    const float scale = (float)((double)prm.gamma[c] * rstd);
    prm.scale[c] = scale;
    prm.shift[c] = (float)((double)prm.beta[c] - mean * (double)scale);
    prm.rstd[c] = (float)rstd;
    if (prm.running_mean) {                                              // moving averages (Keras momentum 0.99 == torch 0.01; unbiased variance)
        const double unb = prm.count > 1.0 ? var * prm.count / (prm.count - 1.0) : var;
        prm.running_mean[c] = (float)((1.0 - prm.momentum) * prm.running_mean[c] + prm.momentum * mean);
        prm.running_var[c] = (float)((1.0 - prm.momentum) * prm.running_var[c] + prm.momentum * unb);
    }
}

struct BnrApplyParams {
    const float *g;   int64_t sg;
    const float *g2;  int64_t sg2;   // nullable
    const float *x;   int64_t sx;
    float *dst;       int64_t sd;
    const float *init; int64_t si;   // nullable: dst = init + value (the first contribution to a gradient slice that already has an upstream part)
    const float *scale, *shift, *mean, *rstd, *g_beta, *g_gamma;
    float inv_n;
    int relu, accumulate;
    uint64_t n;                      // 16-byte vectors
    uint32_t vpp;
    FastDiv div_vpp;
};

// d x = scale * (gm - d beta / n - xhat * d gamma / n), written to or added into dst
__global__ void __launch_bounds__(256) bnr_apply_kernel(const __grid_constant__ BnrApplyParams prm) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= prm.n) return;
    uint32_t p, v;
    prm.div_vpp.divmod((uint32_t)i, p, v);
    const uint32_t c = 4 * v;
    const float4 gv = __ldg(reinterpret_cast<const float4 *>(prm.g + (int64_t)p * prm.sg + c));
    const float4 xv = __ldg(reinterpret_cast<const float4 *>(prm.x + (int64_t)p * prm.sx + c));
    float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), dv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (prm.g2) hv = __ldg(reinterpret_cast<const float4 *>(prm.g2 + (int64_t)p * prm.sg2 + c));
    float4 *dptr = reinterpret_cast<float4 *>(prm.dst + (int64_t)p * prm.sd + c);
    if (prm.accumulate) dv = *dptr;
    else if (prm.init) dv = __ldg(reinterpret_cast<const float4 *>(prm.init + (int64_t)p * prm.si + c));
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(prm.scale + c)), sh = __ldg(reinterpret_cast<const float4 *>(prm.shift + c));
    const float4 mu = __ldg(reinterpret_cast<const float4 *>(prm.mean + c)), rs = __ldg(reinterpret_cast<const float4 *>(prm.rstd + c));
    const float4 gb = __ldg(reinterpret_cast<const float4 *>(prm.g_beta + c)), gg = __ldg(reinterpret_cast<const float4 *>(prm.g_gamma + c));
    const float g[4] = {gv.x, gv.y, gv.z, gv.w}, x[4] = {xv.x, xv.y, xv.z, xv.w}, h[4] = {hv.x, hv.y, hv.z, hv.w};
    const float s[4] = {sc.x, sc.y, sc.z, sc.w}, t[4] = {sh.x, sh.y, sh.z, sh.w}, m[4] = {mu.x, mu.y, mu.z, mu.w}, r[4] = {rs.x, rs.y, rs.z, rs.w};
    const float b1[4] = {gb.x, gb.y, gb.z, gb.w}, b2[4] = {gg.x, gg.y, gg.z, gg.w};
    float d[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool on = !prm.relu || fmaf(x[e], s[e], t[e]) > 0.0f;
        const float gm = (on ? g[e] : 0.0f) + h[e];
        const float xhat = (x[e] - m[e]) * r[e];
        d[e] += s[e] * (gm - b1[e] * prm.inv_n - xhat * (b2[e] * prm.inv_n));
    }
    *dptr = make_float4(d[0], d[1], d[2], d[3]);
}

}  // namespace btslpg
