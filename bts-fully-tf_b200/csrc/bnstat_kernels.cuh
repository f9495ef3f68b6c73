// bnstat_kernels.cuh -- the statistics passes of the TRAINING-mode conv block glue (sm_100a).
//
// bts_decoder.py:30-44 (conv_block): upconv = Conv2D(nf, 3, activation='elu')(upsample) ; upconv = BatchNormalization(momentum=0.99,
// epsilon=1.1e-5)(upconv, training) ; concat = Concatenate([upconv, skip(, lpg)]).  The framework runs ELU, BatchNorm and cat as three
// passes forward (read + write each) and three backward.  Here, with `raw` the convolution's linear output:
//   forward   bn_elu_stats      one READ of raw: per-channel sums of elu(raw) and elu(raw)^2 -> mean, biased variance, the folded
//                               affine (scale = gamma / sqrt(var + eps), shift = beta - mean * scale) and the moving averages;
//             concat_fwd        (existing kernel) elu + affine + concat in one read + one write
//   backward  bn_elu_bwd_stats  one read of the concat gradient's first CA channels and of the saved concat output:
//                               d beta = sum g, d gamma = sum g * xhat with xhat = (y - beta) / gamma recovered from the output
//             concat_bwd        (existing kernel, `bn` pack) g_raw = scale * (g - mean(g) - xhat * mean(g * xhat)) * elu'
// Reductions are deterministic: a thread always owns the same four channels, fixed-order sums thread -> CTA (shared memory, float64)
// -> grid (per-CTA rows in the workspace, summed in a fixed order by bn_finalize_kernel; no atomics at all).
//
// `pack`: float32 [8][C] device buffer shared by the four kernels of a block:
//   [0] scale  [1] shift  [2] mean  [3] std = sqrt(var + eps)  [4] 1/gamma  [5] beta  [6] c1 = mean(g)  [7] c2 = mean(g * xhat)
#pragma once

#include "common.cuh"
#include "concat_kernels.cuh"   // elu_fwd

namespace btslpg {

constexpr int kBnThreads = 256;
constexpr int kBnMaxBlocks = 148 * 4;
constexpr int kBnHeaderBytes = 256;

struct BnStatParams {
    const float *a;            // forward: raw (npix, C); backward: g_out (npix, stride)
    const float *y;            // backward: saved concat output (npix, stride)
    uint64_t npix;
    uint32_t C, stride;        // channels reduced; elements between consecutive pixels
    int act;                   // forward: 1 = statistics of elu(a), 0 = of a
    const float *gamma, *beta;
    float *running_mean, *running_var;   // nullable
    float momentum, eps;       // torch convention: running = (1 - momentum) * running + momentum * batch
    float *pack;               // [8][C]
    float *g_gamma, *g_beta;   // backward outputs [C]
    double *partial;           // [gridDim.x][2][C]
};

// BWD = false: sums of v and v^2 (v = elu(a) or a);  BWD = true: sums of g and g * xhat
template <bool BWD> __global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const __grid_constant__ BnStatParams prm) {
    extern __shared__ __align__(16) unsigned char bn_smem[];
    double *red = reinterpret_cast<double *>(bn_smem);                  // [kBnThreads][8]
    const uint32_t C = prm.C;
    const uint32_t vec_per_px = C / 4;                                   // 16-byte vectors per pixel
    // thread t always owns channels 4 * (t % vec_per_px) .. + 3 (kBnThreads is a multiple of vec_per_px or vice versa: C is a power-of-two multiple of 4 up to 1024)
    const uint32_t lanes = kBnThreads >= vec_per_px ? vec_per_px : kBnThreads;          // distinct channel groups per pass
    const uint32_t cg0 = threadIdx.x % lanes;
    const uint32_t px_per_pass = kBnThreads >= vec_per_px ? kBnThreads / vec_per_px : 1;
    const uint32_t groups = vec_per_px / lanes;                          // channel-group passes per pixel when C > 4 * kBnThreads (never for C <= 1024)
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    float ig[4], bt[4];
    const uint32_t cbase = 4 * cg0;
    if (BWD) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { ig[e] = prm.pack[4 * C + cbase + e]; bt[e] = prm.pack[5 * C + cbase + e]; }
    }
    (void)groups;
    const uint64_t px0 = (uint64_t)blockIdx.x * px_per_pass + threadIdx.x / lanes;
    const uint64_t step = (uint64_t)gridDim.x * px_per_pass;
    double d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
    uint32_t run = 0;
    constexpr int U = 4;                                                 // pixels per thread and trip: all loads first (memory-level parallelism)
    for (uint64_t p = px0; p < prm.npix; p += U * step) {
        float4 v[U], yv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t q = p + u * step;
            if (q < prm.npix) {
                v[u] = __ldg(reinterpret_cast<const float4 *>(prm.a + q * prm.stride + cbase));
                if (BWD) yv[u] = __ldg(reinterpret_cast<const float4 *>(prm.y + q * prm.stride + cbase));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (p + u * step >= prm.npix) break;
            const float x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            if (BWD) {
                const float yy[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float xhat = (yy[e] - bt[e]) * ig[e];
                    s1[e] += x[e];
                    s2[e] = fmaf(x[e], xhat, s2[e]);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float w = prm.act ? elu_fwd(x[e]) : x[e];
                    s1[e] += w;
                    s2[e] = fmaf(w, w, s2[e]);
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
    for (int e = 0; e < 4; ++e) { d1[e] += (double)s1[e]; d2[e] += (double)s2[e]; }
#pragma unroll
    for (int e = 0; e < 4; ++e) { red[threadIdx.x * 8 + e] = d1[e]; red[threadIdx.x * 8 + 4 + e] = d2[e]; }
    __syncthreads();
    // threads -> CTA: channel c = 4 * cg + e is owned by threads cg, cg + lanes, ... (fixed order)
    for (uint32_t i = threadIdx.x; i < 2 * C; i += kBnThreads) {
        const uint32_t which = i / C, c = i % C, cg = c / 4, e = c % 4;
        double s = 0.0;
        if (cg < lanes)
            for (uint32_t t = cg; t < kBnThreads; t += lanes) s += red[t * 8 + which * 4 + e];
        prm.partial[(size_t)blockIdx.x * 2 * C + i] = s;
    }
}

// CTAs -> result.  A second, wide launch instead of a last-CTA tail: up to 592 rows x 2C columns of float64 partials would be a
// chain of dependent L2 round trips for one CTA (tens of microseconds at C = 512); here 8 channels per CTA, 32 row slices per
// channel, four loads in flight per thread, slices combined in slice order through shared memory: fixed order, no atomics.
template <bool BWD> __global__ void __launch_bounds__(256) bn_finalize_kernel(const __grid_constant__ BnStatParams prm, uint32_t nrows) {
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
    const double n = (double)prm.npix;
    if (BWD) {
        prm.g_beta[c] = (float)a1;                                   // d loss / d beta
        prm.g_gamma[c] = (float)a2;                                  // d loss / d gamma
        prm.pack[6 * C + c] = (float)(a1 / n);
        prm.pack[7 * C + c] = (float)(a2 / n);
    } else {
        const double mean = a1 / n;
        double var = a2 / n - mean * mean;                           // biased: what training-mode normalisation uses
        if (var < 0.0) var = 0.0;
        const double std = sqrt(var + (double)prm.eps);
        const float gamma = prm.gamma[c], beta = prm.beta[c];
        const float scale = (float)((double)gamma / std);
        prm.pack[0 * C + c] = scale;
        prm.pack[1 * C + c] = (float)((double)beta - mean * (double)scale);
        prm.pack[2 * C + c] = (float)mean;
        prm.pack[3 * C + c] = (float)std;
        prm.pack[4 * C + c] = 1.0f / gamma;
        prm.pack[5 * C + c] = beta;
        if (prm.running_mean) {                                      // moving averages (Keras momentum 0.99 == torch momentum 0.01; unbiased variance)
            const double unb = n > 1.0 ? var * n / (n - 1.0) : var;
            prm.running_mean[c] = (float)((1.0 - prm.momentum) * prm.running_mean[c] + prm.momentum * mean);
            prm.running_var[c] = (float)((1.0 - prm.momentum) * prm.running_var[c] + prm.momentum * unb);
        }
    }
}

}  // namespace btslpg
