// tail_kernels.cuh -- the full-resolution tail of the BTS decoder after the last convolution (sm_100a):
// SURVEY 8(f) rows N2 and N4, the first two "next" rows around the LPG hot path.
//
//   silog forward   depth_est = sigmoid(logit) * max_depth            bts_decoder.py:102-103
//                   loss = si_log_loss(y_true, depth_est)             bts.py:27-41          (one pass)
//   silog backward  d loss / d logit                                  (TF autodiff of the two above)
//   eval metrics    silog, abs_rel, log10, rmse, sq_rel, rmse_log, d1, d2, d3
//                                                                     custom_eval_metrics.py:24-88 (one pass
//                   instead of nine separately masked reductions, each re-running pre_eval)
//
// All three are HBM-streaming passes over (B,H,W,1) maps with a handful of scalar sums.  The sums are
// deterministic: fixed-order partial sums per thread (float32, a few dozen terms), a fixed xor-shuffle
// tree per warp, then float64 from the warp level up -- warps -> CTA through shared memory, CTAs ->
// result through a workspace that the last CTA to finish adds in a fixed order (integer completion
// counter only, no float atomics).  Results do not depend on scheduling.
#pragma once

#include "common.cuh"

namespace btslpg {

#ifndef BTSLPG_TAIL_THREADS
#define BTSLPG_TAIL_THREADS 1024     // one fat CTA per SM: 148 completion-counter atomics and partial rows instead of 740 (256: fwd 25.3, metrics 28.2 us; 1024: 24.1, 22.1)
#endif
constexpr int kTailThreads = BTSLPG_TAIL_THREADS;
constexpr int kTailMaxBlocks = 148 * 8;      // persistent grid: at most 8 CTAs per SM
constexpr int kTailMaxSums = 10;             // the metrics kernel carries 10 sums, the loss 3
constexpr int kTailHeaderBytes = 256;        // completion counter
constexpr int kTailStatsDoubles = 16;        // results kept for the backward pass (n, mean d, variance term ...)
constexpr size_t kTailWorkspaceBytes =
    kTailHeaderBytes + kTailStatsDoubles * sizeof(double) + (size_t)kTailMaxBlocks * kTailMaxSums * sizeof(double);

struct TailWorkspace {
    unsigned int *counter;
    double *stats;
    double *partial;
    __host__ __device__ explicit TailWorkspace(void *ws)
        : counter(reinterpret_cast<unsigned int *>(ws)),
          stats(reinterpret_cast<double *>(static_cast<char *>(ws) + kTailHeaderBytes)),
          partial(reinterpret_cast<double *>(static_cast<char *>(ws) + kTailHeaderBytes) + kTailStatsDoubles) {}
};

// Sum NV per-thread values over the whole grid.  Returns true in every thread of the LAST CTA to finish,
// with the totals in tot[] (shared memory); false elsewhere.  Fixed order at every level.
template <int NV> __device__ __forceinline__ bool tail_grid_sum(const float (&v)[NV], const TailWorkspace &ws, double *tot /* shared [NV] */) {
    __shared__ double red[kTailThreads / 32][NV];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float w[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        w[k] = v[k];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) w[k] += __shfl_xor_sync(0xffffffffu, w[k], m);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[wid][k] = (double)w[k];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < kTailThreads / 32; ++q) s += red[q][threadIdx.x];
        ws.partial[(size_t)blockIdx.x * NV + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(ws.counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    // one warp per sum: lanes walk the CTA partials (independent loads), then a fixed xor tree in float64
    for (int k = wid; k < NV; k += kTailThreads / 32) {
        double s = 0.0;
        for (uint32_t blk = lane; blk < gridDim.x; blk += 32) s += __ldcg(ws.partial + (size_t)blk * NV + k);
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        if (lane == 0) tot[k] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) *ws.counter = 0u;     // leave the header zero for the next launch
    return true;
}

// Transcendentals.  These passes move 8-12 bytes per element, so they only stay on the HBM roofline if an
// element costs ~20 instructions: log, exp and 1/x come from the SFU (MUFU.LG2 / EX2 / RCP, each within
// ~2 ulp, i.e. <= 2.4e-7 relative / 1.2e-7 absolute on a logarithm) instead of the ~20-instruction libm
// forms -- two orders of magnitude inside the 1e-5 parity budget, and of the size of TensorFlow's own
// kernel-to-kernel differences.  Inputs are never subnormal here (y + 1e-7, clipped predictions), so the
// .ftz forms are exact about that.
__device__ __forceinline__ float lg2_sfu(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_sfu(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
constexpr float kLn2 = 0.693147180559945309f;
constexpr float kLog2e = 1.442695040888963407f;
// log(a) - log(b)
__device__ __forceinline__ float log_ratio(float a, float b) { return (lg2_sfu(a) - lg2_sfu(b)) * kLn2; }
// bts_decoder.py:102 activation='sigmoid': 1 / (1 + exp(-z));  z -> -inf gives exp = +inf, 1/inf = 0
__device__ __forceinline__ float tail_sigmoid(float z) { return rcp_approx(1.0f + ex2_sfu(-z * kLog2e)); }

template <typename T> struct TailVec {
    static constexpr int N = 16 / (int)sizeof(T);      // elements per 16-byte access: 4 float32 / 8 bfloat16
};

// ------------------------------------------------------------------------------------------------
// silog forward.  Per element: s = sigmoid(z); y = s*max_depth (stored); if y_true > th:
// d = log(y_true + eps) - log(y + eps) (bts.py:37), sums n, d, d^2.  Finalisation by the last CTA:
// loss = sqrt(mean(d^2) - 0.85*mean(d)^2) * 10 (bts.py:38); an empty mask gives NaN as in the
// reference (mean of an empty tensor).  Algorithmic bytes per element: 3 * sizeof(T) fused (read logit
// and y_true, write depth_est), 2 * sizeof(T) for the loss alone on a given depth_est.
// ------------------------------------------------------------------------------------------------
template <typename T> struct SilogFwdParams {
    const T *logit;      // nullable: the loss alone, on a given depth_est (the reference's si_log_loss(y_true, y_pred))
    const T *y_true;     // nullable: depth only
    T *depth;            // written when logit is given, read otherwise
    float *loss;         // device scalar
    void *workspace;
    uint64_t n;
    float max_depth, threshold;
};

__device__ __forceinline__ void silog_accum(float yt, float y, float th, float (&acc)[3]) {
    const bool valid = yt > th;                                                       // bts.py:32 K.greater(y_true, gt_th)
    const float d = valid ? log_ratio(yt + BTSLPG_EPS_F, y + BTSLPG_EPS_F) : 0.0f;   // bts.py:37 (predicated, no divergence)
    acc[0] += valid ? 1.0f : 0.0f;
    acc[1] += d;
    acc[2] = fmaf(d, d, acc[2]);
}

template <typename T> __global__ void __launch_bounds__(kTailThreads) silog_fwd_kernel(const __grid_constant__ SilogFwdParams<T> prm) {
    constexpr int N = TailVec<T>::N;
    __shared__ double tot[3];
    const uint64_t nvec = prm.n / N;
    const uint64_t stride = (uint64_t)gridDim.x * kTailThreads;
    float acc[3] = {0.0f, 0.0f, 0.0f};
    // two vectors per thread and iteration: all four loads are issued before the first use
    for (uint64_t i = (uint64_t)blockIdx.x * kTailThreads + threadIdx.x; i < nvec; i += 2 * stride) {
        const uint64_t i2 = i + stride;
        const bool two = i2 < nvec;
        float z[2][N], yt[2][N], y[2][N];
        if (prm.logit) {
            load_elems<T, N, 4>(prm.logit + i * N, z[0]);
            if (two) load_elems<T, N, 4>(prm.logit + i2 * N, z[1]);
        } else {
            load_elems<T, N, 4>(prm.depth + i * N, y[0]);
            if (two) load_elems<T, N, 4>(prm.depth + i2 * N, y[1]);
        }
        if (prm.y_true) {
            load_elems<T, N, 4>(prm.y_true + i * N, yt[0]);
            if (two) load_elems<T, N, 4>(prm.y_true + i2 * N, yt[1]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            const uint64_t iu = u ? i2 : i;
            if (prm.logit) {
#pragma unroll
                for (int e = 0; e < N; ++e) y[u][e] = tail_sigmoid(z[u][e]) * prm.max_depth;   // bts_decoder.py:102-103
                store_elems<T, N, 4>(prm.depth + iu * N, y[u]);
            }
            if (prm.y_true) {
                // the loss sees depth_est as stored (rounded to T), exactly what a downstream loss op would read
#pragma unroll
                for (int e = 0; e < N; ++e)
                    silog_accum(yt[u][e], sizeof(T) == 2 ? __bfloat162float(__float2bfloat16_rn(y[u][e])) : y[u][e], prm.threshold, acc);
            }
        }
    }
    // ragged tail (n % N elements), one thread
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (uint64_t i = nvec * N; i < prm.n; ++i) {
            float y;
            if (prm.logit) {
                y = tail_sigmoid(load1(prm.logit + i)) * prm.max_depth;
                store1(prm.depth + i, y);
            } else {
                y = load1(prm.depth + i);
            }
            if (prm.y_true) silog_accum(load1(prm.y_true + i), sizeof(T) == 2 ? __bfloat162float(__float2bfloat16_rn(y)) : y, prm.threshold, acc);
        }
    }
    if (!prm.y_true) return;      // uniform across the grid
    const TailWorkspace ws(prm.workspace);
    if (tail_grid_sum<3>(acc, ws, tot) && threadIdx.x == 0) {
        const double n = tot[0], m1 = tot[1] / n, m2 = tot[2] / n;
        const double var = m2 - 0.85 * m1 * m1;
        ws.stats[0] = n;
        ws.stats[1] = m1;
        ws.stats[2] = var;
        *prm.loss = (float)(sqrt(var) * 10.0);
    }
}

// ------------------------------------------------------------------------------------------------
// silog backward: with V = mean(d^2) - 0.85 mean(d)^2, loss = 10 sqrt(V):
//   d loss / d d_i   = 10 (d_i - 0.85 mean d) / (n sqrt V)          (masked elements, else 0)
//   d d_i / d y_i    = -1 / (y_i + eps)
//   d y_i / d z_i    = max_depth s_i (1 - s_i) = y_i (1 - y_i / max_depth)
// Reads depth_est (saved output) and y_true, writes d loss / d logit.  3 * sizeof(T) bytes per element.
// ------------------------------------------------------------------------------------------------
template <typename T> struct SilogBwdParams {
    const T *depth;
    const T *y_true;
    const float *g_loss;   // device scalar, nullable (= 1)
    const void *workspace; // as left by the forward
    T *g_logit;            // d loss / d logit (wrt_logit) or d loss / d depth_est
    uint64_t n;
    float max_depth, threshold;
    int wrt_logit;
};

__device__ __forceinline__ float silog_grad(float yt, float y, float th, float c1, float m1s, float inv_md) {
    const float a = y + BTSLPG_EPS_F;
    const float d = log_ratio(yt + BTSLPG_EPS_F, a);
    const float gd = c1 * (d - m1s);                    // d loss / d d_i
    const float gy = -gd * rcp_approx(a);               // d loss / d y_i
    const float g = inv_md > 0.0f ? gy * (y * (1.0f - y * inv_md)) : gy;   // d loss / d z_i when differentiating through the sigmoid
    return yt > th ? g : 0.0f;
}

template <typename T> __global__ void __launch_bounds__(kTailThreads) silog_bwd_kernel(const __grid_constant__ SilogBwdParams<T> prm) {
    constexpr int N = TailVec<T>::N;
    const TailWorkspace ws(const_cast<void *>(prm.workspace));
    const double n = ws.stats[0], m1 = ws.stats[1], var = ws.stats[2];
    const float gl = prm.g_loss ? *prm.g_loss : 1.0f;
    const float c1 = (float)((double)gl * 10.0 / (n * sqrt(var)));
    const float m1s = (float)(0.85 * m1);
    const float inv_md = prm.wrt_logit ? 1.0f / prm.max_depth : 0.0f;
    const uint64_t nvec = prm.n / N;
    const uint64_t stride = (uint64_t)gridDim.x * kTailThreads;
    for (uint64_t i = (uint64_t)blockIdx.x * kTailThreads + threadIdx.x; i < nvec; i += 2 * stride) {
        const uint64_t i2 = i + stride;
        const bool two = i2 < nvec;
        float y[2][N], yt[2][N], g[N];
        load_elems<T, N, 4>(prm.depth + i * N, y[0]);
        load_elems<T, N, 4>(prm.y_true + i * N, yt[0]);
        if (two) {
            load_elems<T, N, 4>(prm.depth + i2 * N, y[1]);
            load_elems<T, N, 4>(prm.y_true + i2 * N, yt[1]);
        }
#pragma unroll
        for (int e = 0; e < N; ++e) g[e] = silog_grad(yt[0][e], y[0][e], prm.threshold, c1, m1s, inv_md);
        store_elems<T, N, 4>(prm.g_logit + i * N, g);
        if (two) {
#pragma unroll
            for (int e = 0; e < N; ++e) g[e] = silog_grad(yt[1][e], y[1][e], prm.threshold, c1, m1s, inv_md);
            store_elems<T, N, 4>(prm.g_logit + i2 * N, g);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (uint64_t i = nvec * N; i < prm.n; ++i)
            store1(prm.g_logit + i, silog_grad(load1(prm.y_true + i), load1(prm.depth + i), prm.threshold, c1, m1s, inv_md));
    }
}

// ------------------------------------------------------------------------------------------------
// Eval metrics (custom_eval_metrics.py:24-88), all nine from ONE pass over (y_true, y_pred):
//   pre_eval (:39-42): mask = min < y_true < max ; pred = clip(where(isfinite(pred), pred, max), min, max)
//   sums over masked elements: n, log-diff d (:64,:79,:85 -> d, d^2, |d|), (gt-pred)^2, |gt-pred|/gt,
//   (gt-pred)^2/gt, [max(gt/pred, pred/gt) < 1.25^k] for k = 1, 2, 3.
// Output order = the reference's list (:88): silog, abs_rel, log10, rmse, sq_rel, rmse_log, d1, d2, d3,
// followed by n.  2 * sizeof(T) bytes per element.
// ------------------------------------------------------------------------------------------------
template <typename T> struct MetricsParams {
    const T *y_true;
    const T *y_pred;
    float *out;          // [10] device
    void *workspace;
    uint64_t n;
    float min_depth, max_depth;
    uint16_t *png;       // bts_predict.py:140-141 output, nullable (PNG template flag)
    float png_max_depth;
};

// bts_predict.py:140-141: `pred_depth * 65536 / args.max_depth` in float32, then numpy's `.astype(np.uint16)`: the C cast of
// x86-64 numpy (float -> int32 by truncation, low 16 bits kept: a prediction that rounds to max_depth wraps to 0, NaN and
// out-of-range values give 0 -- the reference does not clamp, so neither does this).
__device__ __forceinline__ uint16_t png16(float depth, float max_depth) {
    const float v = __fdiv_rn(depth * 65536.0f, max_depth);
    const int i = (fabsf(v) < 2147483648.0f) ? __float2int_rz(v) : (int)0x80000000;      // cvttss2si: "integer indefinite" when out of range / NaN
    return (uint16_t)(unsigned)i;
}
template <int N> __device__ __forceinline__ void png16_store(uint16_t *dst, const float (&pr)[N], float max_depth) {
    uint32_t w[N / 2];
#pragma unroll
    for (int e = 0; e < N / 2; ++e) w[e] = (uint32_t)png16(pr[2 * e], max_depth) | ((uint32_t)png16(pr[2 * e + 1], max_depth) << 16);
    if constexpr (N == 4) *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
    else *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ void metrics_accum(float gt, float pr, float lo_, float hi_, float (&acc)[10]) {
    const bool valid = gt < hi_ && gt > lo_;                                 // :39
    float p = (fabsf(pr) <= 3.402823466e+38f) ? pr : hi_;                    // :40 (NaN and +-inf -> max_depth_eval)
    p = fminf(fmaxf(p, lo_), hi_);                                           // :41
    p = valid ? p : 1.0f;                                                    // masked-out pixels: harmless operands, zero weight
    const float g = valid ? gt : 1.0f;
    const float w = valid ? 1.0f : 0.0f;
    const float rg = rcp_approx(g), rp = rcp_approx(p);
    const float d = w * log_ratio(g, p);                                     // :64, :79, :85
    const float diff = g - p;
    const float sq = diff * diff;
    const float ratio = fmaxf(g * rp, p * rg);                               // :47
    acc[0] += w;
    acc[1] += d;
    acc[2] = fmaf(d, d, acc[2]);
    acc[3] += fabsf(d);
    acc[4] = fmaf(w, sq, acc[4]);                                            // :60
    acc[5] = fmaf(w * fabsf(diff), rg, acc[5]);                              // :70
    acc[6] = fmaf(w * sq, rg, acc[6]);                                       // :74
    acc[7] += ratio < 1.25f ? w : 0.0f;                                      // :47
    acc[8] += ratio < 1.5625f ? w : 0.0f;                                    // :51  1.25 ** 2
    acc[9] += ratio < 1.953125f ? w : 0.0f;                                  // :55  1.25 ** 3
}

// PNG: additionally write the uint16 depth image of bts_predict.py:140-141 from the same read of y_pred (N4's second half);
// METRICS = false (no ground truth, as in bts_predict.py): the scaling pass alone.
template <typename T, bool PNG, bool METRICS> __global__ void __launch_bounds__(kTailThreads) eval_metrics_kernel(const __grid_constant__ MetricsParams<T> prm) {
    constexpr int N = TailVec<T>::N;
    __shared__ double tot[10];
    const uint64_t nvec = prm.n / N;
    const uint64_t stride = (uint64_t)gridDim.x * kTailThreads;
    float acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.0f;
    for (uint64_t i = (uint64_t)blockIdx.x * kTailThreads + threadIdx.x; i < nvec; i += 2 * stride) {
        const uint64_t i2 = i + stride;
        const bool two = i2 < nvec;
        float gt[2][N], pr[2][N];
        if constexpr (METRICS) load_elems<T, N, 4>(prm.y_true + i * N, gt[0]);
        load_elems<T, N, 4>(prm.y_pred + i * N, pr[0]);
        if (two) {
            if constexpr (METRICS) load_elems<T, N, 4>(prm.y_true + i2 * N, gt[1]);
            load_elems<T, N, 4>(prm.y_pred + i2 * N, pr[1]);
        }
        if constexpr (PNG) png16_store<N>(prm.png + i * N, pr[0], prm.png_max_depth);
        if constexpr (METRICS) {
#pragma unroll
            for (int e = 0; e < N; ++e) metrics_accum(gt[0][e], pr[0][e], prm.min_depth, prm.max_depth, acc);
        }
        if (two) {
            if constexpr (PNG) png16_store<N>(prm.png + i2 * N, pr[1], prm.png_max_depth);
            if constexpr (METRICS) {
#pragma unroll
                for (int e = 0; e < N; ++e) metrics_accum(gt[1][e], pr[1][e], prm.min_depth, prm.max_depth, acc);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (uint64_t i = nvec * N; i < prm.n; ++i) {
            const float pv = load1(prm.y_pred + i);
            if constexpr (PNG) prm.png[i] = png16(pv, prm.png_max_depth);
            if constexpr (METRICS) metrics_accum(load1(prm.y_true + i), pv, prm.min_depth, prm.max_depth, acc);
        }
    }
    if constexpr (!METRICS) return;
    const TailWorkspace ws(prm.workspace);
    if (tail_grid_sum<10>(acc, ws, tot) && threadIdx.x == 0) {
        const double n = tot[0], m1 = tot[1] / n, m2 = tot[2] / n;
        prm.out[0] = (float)(sqrt(m2 - m1 * m1) * 100.0);        // silog   :80
        prm.out[1] = (float)(tot[5] / n);                        // abs_rel :70
        prm.out[2] = (float)(tot[3] / n / log(10.0));            // log10   :86
        prm.out[3] = (float)sqrt(tot[4] / n);                    // rmse    :60
        prm.out[4] = (float)(tot[6] / n);                        // sq_rel  :74
        prm.out[5] = (float)sqrt(m2);                            // rmse_log :65
        prm.out[6] = (float)(tot[7] / n);                        // d1
        prm.out[7] = (float)(tot[8] / n);                        // d2
        prm.out[8] = (float)(tot[9] / n);                        // d3
        prm.out[9] = (float)n;
    }
}

}  // namespace btslpg
