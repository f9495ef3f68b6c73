// depthconv_kernels.cuh -- backward of the decoder's last convolution (sm_100a), SURVEY 8(f) N1:
//     depth_est_scaled = Conv2D(1, kernel_size=3, padding='same', use_bias=False)(iconv1)        bts_decoder.py:102
// i.e. y[p] = sum_{t, c} x[p + t][c] * w[t][c] with t over the 3x3 taps and C = F/16 input channels, ONE output channel.
//
// cuDNN runs the forward and the data gradient of this layer near their HBM floors, but its weight gradient is a
// tensor-core GEMM with a 288-element output and a 10-million-term reduction: 2.05 ms at B = 32, 480x640 for 0.2 ms of
// traffic, plus two layout conversions (1.1 ms) because of the single output channel -- 3.7 ms for a backward whose
// floor is 0.4 ms.  With one output channel the arithmetic is tiny (2 x 9 FMA per input element), so plain FP32 FMAs
// keep up with HBM and both gradients come from ONE pass:
//     g_x[q][c] = sum_t g[q - t] * w[t][c]                (written once)
//     g_w[t][c] = sum_q g[q - t] * x[q][c]                (accumulated in registers, reduced in a fixed order)
// A thread owns 4 channels of one pixel at a time (C/4 lanes per pixel: a warp instruction touches 128-512
// contiguous bytes of x and of g_x); the nine upstream gradients a pixel needs come from a 3-row strip of g staged in
// shared memory with zero halos (padding='same').  The 36 weights of the thread's channels and its 36 partial sums
// of g_w live in registers for the whole kernel.  Exact float32 arithmetic (the library conv is TF32).
// Algorithmic bytes per pixel: (2 C + 1) * sizeof(T)  (x read, g_x written, g read).
#pragma once

#include "common.cuh"

namespace btslpg {

constexpr int kDcThreads = 256;
constexpr int kDcTileW = 128;            // pixels of one image row per work item
constexpr int kDcHeaderBytes = 256;
constexpr int kDcMaxBlocks = 148 * 4;

template <typename T> struct DepthConvBwdParams {
    const T *x;          // (B,H,W,C)
    const T *g;          // (B,H,W) upstream gradient of the single output channel
    const float *w;      // [9][C]  (Keras HWIO (3,3,C,1) flattened)
    T *g_x;              // (B,H,W,C), nullable
    float *g_w;          // [9][C], nullable
    float *partial;      // [gridDim.x][9*C]
    unsigned int *counter;
    uint32_t B, H, W, col_blocks, items;
    FastDiv div_cb, div_h;
};

// the last CTA adds the per-CTA partial rows in a fixed order: float4 columns x slices of the CTA range, 8 loads in flight
__device__ __forceinline__ void dc_reduce_partials(const float *partial, uint32_t nblk, uint32_t ncol4, float *out, float4 *comb /* [S][ncol4] */) {
    const uint32_t S = kDcThreads / ncol4;
    const uint32_t col = threadIdx.x % ncol4, sl = threadIdx.x / ncol4;
    if (sl < S) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 *p4 = reinterpret_cast<const float4 *>(partial) + col;
        uint32_t blk = sl;
        for (; blk + 7 * S < nblk; blk += 8 * S) {
            float4 t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = __ldcg(p4 + (size_t)(blk + u * S) * ncol4);
#pragma unroll
            for (int u = 0; u < 8; ++u) { v.x += t[u].x; v.y += t[u].y; v.z += t[u].z; v.w += t[u].w; }
        }
        for (; blk < nblk; blk += S) {
            const float4 t = __ldcg(p4 + (size_t)blk * ncol4);
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        comb[sl * ncol4 + col] = v;
    }
    __syncthreads();
    if (threadIdx.x < ncol4) {
        float4 v = comb[threadIdx.x];
        for (uint32_t q = 1; q < S; ++q) {
            const float4 t = comb[q * ncol4 + threadIdx.x];
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        float *o = out + 4 * threadIdx.x;
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
}

template <typename T, int C> __global__ void __launch_bounds__(kDcThreads, 2) depthconv_bwd_kernel(const __grid_constant__ DepthConvBwdParams<T> prm) {
    constexpr int LPP = C / 4;                 // lanes per pixel
    constexpr int PPW = 32 / LPP;              // pixels per warp pass
    constexpr int NW = kDcThreads / 32;
    constexpr int NCOL4 = 9 * C / 4;           // float4 columns of g_w: 72 (C = 32) or 36 (C = 16)
    __shared__ float gs[3][kDcTileW + 2];
    __shared__ __align__(16) float red[NW][9 * C];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int cg = lane % LPP, pl = lane / LPP;

    float wr[9][4], acc[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            wr[t][e] = __ldg(prm.w + t * C + 4 * cg + e);
            acc[t][e] = 0.0f;
        }

    for (uint32_t item = blockIdx.x; item < prm.items; item += gridDim.x) {
        uint32_t rowi, xb, b, y;
        prm.div_cb.divmod(item, rowi, xb);     // rowi = b * H + y
        prm.div_h.divmod(rowi, b, y);
        const uint32_t x0 = xb * kDcTileW;
        const uint32_t npx = min((uint32_t)kDcTileW, prm.W - x0);
        // 3-row strip of g with a one-pixel halo, zeros outside the image (padding='same')
        for (uint32_t i = threadIdx.x; i < 3 * (kDcTileW + 2); i += kDcThreads) {
            const uint32_t r = i / (kDcTileW + 2), j = i % (kDcTileW + 2);
            const int yy = (int)y + (int)r - 1, xx = (int)x0 + (int)j - 1;
            float v = 0.0f;
            if (yy >= 0 && yy < (int)prm.H && xx >= 0 && xx < (int)prm.W && j <= npx + 1)
                v = load1(prm.g + ((size_t)b * prm.H + yy) * prm.W + xx);
            gs[r][j] = v;
        }
        __syncthreads();
        const size_t row0 = ((size_t)rowi * prm.W + x0);
        // U pixels per thread and iteration: all U loads of x are issued before the first use (the kernel holds ~110
        // registers, i.e. 16 resident warps per SM -- one 16-byte load per thread in flight is latency-bound at 2.6 TB/s)
        constexpr int U = kDcTileW / (NW * PPW) > 4 ? 4 : kDcTileW / (NW * PPW);
        for (uint32_t base = wid * PPW + pl; base < npx; base += U * NW * PPW) {
            float xv[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t px = base + u * NW * PPW;
                if (px < npx) load_elems<T, 4>(prm.x + (row0 + px) * C + 4 * cg, xv[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t px = base + u * NW * PPW;
                if (px >= npx) break;
                float out[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int t = (dy + 1) * 3 + (dx + 1);
                        const float gv = gs[1 - dy][px + 1 - dx];          // g at pixel q - (dy, dx)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            out[e] = fmaf(gv, wr[t][e], out[e]);
                            acc[t][e] = fmaf(gv, xv[u][e], acc[t][e]);
                        }
                    }
                if (prm.g_x) store_elems<T, 4>(prm.g_x + (row0 + px) * C + 4 * cg, out);
            }
        }
        __syncthreads();
    }

    if (!prm.g_w) return;     // uniform across the grid
    // lanes that share a channel group (fixed xor tree over the pixel slots), then warps, then CTAs
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v = acc[t][e];
#pragma unroll
            for (int m = LPP; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
            if (pl == 0) red[wid][t * C + 4 * cg + e] = v;
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * C; i += kDcThreads) {
        float v = 0.0f;
#pragma unroll
        for (int q = 0; q < NW; ++q) v += red[q][i];
        prm.partial[(size_t)blockIdx.x * (9 * C) + i] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(prm.counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        dc_reduce_partials(prm.partial, gridDim.x, NCOL4, prm.g_w, reinterpret_cast<float4 *>(&red[0][0]));
        if (threadIdx.x == 0) *prm.counter = 0u;      // leave the workspace header zero for the next launch
    }
}

}  // namespace btslpg
