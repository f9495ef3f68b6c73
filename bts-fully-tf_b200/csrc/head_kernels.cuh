// head_kernels.cuh -- reduction_{8x8,4x4,2x2} head fused with its LPG layer (sm_100a).
//
// Reference (bts_decoder.py:79-81, 86-88, 93-94):
//     reduction = Conv2D(3, 1x1, sigmoid, use_bias=False)(feat)        # cuDNN 1x1 conv + sigmoid
//     depth     = LocalPlanarGuidance(r)(reduction)                    # see lpg_kernels.cuh
//     depth_ds  = depth[:, ::d, ::d]
// The head reads C = 32..128 channels per coarse pixel and is >= 90 % of the fused path's bytes
// at 1.4 flop/byte: HBM-bound by a wide margin, so the contraction runs on the FP32
// pipe and no tensor-core instruction is used.
//
// Work decomposition (forward and backward): a warp iteration covers 32 consecutive coarse pixels.
//   phase 1  8 lanes share one pixel: lane s reads channels [32m+4s, 32m+4s+4) for m < C/32, so
//            every load instruction fetches four whole 128-byte lines (one per lane group); the
//            [C][3] kernel slice a lane needs lives in registers for the whole kernel.
//   phase 2  a 3-level butterfly (reduce-scatter over the 8 lanes, 21 shuffles per 32 pixels)
//            leaves lane L holding the three pre-activations of pixel L.
//   phase 3  lane L applies the sigmoid, stores the coefficients and expands ITS pixel with the
//            same code as the stand-alone LPG kernel (one coarse pixel per lane, rows of r elements).
// Backward mirrors it: lane L reduces the r x r patch of its pixel (fixed order, in registers),
// forms dz = g_coef*x*(1-x), broadcasts it to the 8 lanes that own the pixel's channels, which
// write g_feat lines and accumulate g_kernel partials.  g_kernel is then reduced lanes -> warps
// (shared memory, fixed order) -> CTAs (workspace, fixed order, last CTA to finish sums): no float
// atomics anywhere, bit-reproducible for a given launch configuration.
#pragma once

#include "lpg_kernels.cuh"
#include "tma_pipe.cuh"

namespace btslpg {

constexpr int kHeadWorkspaceHeader = 256;  // bytes; holds the CTA completion counter
constexpr int kHeadMaxBlocks = 2048;

template <typename T> struct HeadFwdParams {
    const T *feat;        // (npix, C) contiguous NHWC
    const float *kernel;  // [C][3]
    T *coef_out;          // (npix, 3) contiguous
    T *out;
    T *ds;                // nullable
    uint32_t out_sB, out_sH, ds_sB, ds_sH;   // < 2^31, checked on the host
    uint32_t npix, iters;
    FastDiv w, h;
};

template <typename T> struct HeadBwdParams {
    const T *feat;
    const float *kernel;
    const T *coef;        // saved sigmoid output (npix, 3)
    const T *g_full;      // nullable
    const T *g_ds;        // nullable
    uint32_t gf_sB, gf_sH, gd_sB, gd_sH;
    T *g_feat;            // nullable
    float *g_kernel;      // nullable, [C][3]
    T *g_coef_out;        // nullable
    float *partial;       // [gridDim.x][C*3]
    unsigned int *counter;
    uint32_t npix, iters;
    FastDiv w, h;
};

__device__ __forceinline__ float sigmoidf_acc(float z) { return __fdiv_rn(1.0f, 1.0f + expf(-z)); }

template <typename T> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_to<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// reduce-scatter of v[8][3] over the 8 lanes of a pixel group: lane s ends with the total of row s
__device__ __forceinline__ void butterfly8x3(float (&v)[8][3], int s, float (&out)[3]) {
    const bool h4 = s & 4, h2 = s & 2, h1 = s & 1;
    float a[4][3], b[2][3];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float send = h4 ? v[j][k] : v[j + 4][k];
            const float keep = h4 ? v[j + 4][k] : v[j][k];
            a[j][k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float send = h2 ? a[j][k] : a[j + 2][k];
            const float keep = h2 ? a[j + 2][k] : a[j][k];
            b[j][k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float send = h1 ? b[0][k] : b[1][k];
        const float keep = h1 ? b[1][k] : b[0][k];
        out[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
}

template <typename T, int R, int D, int M>
__global__ void __launch_bounds__(256) head_lpg_fwd_kernel(const __grid_constant__ HeadFwdParams<T> prm) {
    constexpr int C = 32 * M;
    constexpr int NDS = D ? R / D : 0;
    const int lane = threadIdx.x & 31, s = lane & 7, gbase = lane & 24;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;

    float wk[M][4][3];
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int k = 0; k < 3; ++k) wk[m][e][k] = __ldg(prm.kernel + (32 * m + 4 * s + e) * 3 + k);

    for (uint32_t iter = warp; iter < prm.iters; iter += nwarps) {
        const uint32_t p0 = iter * 32;
        float acc[8][3];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            acc[it][0] = acc[it][1] = acc[it][2] = 0.0f;
            const uint32_t pix = p0 + gbase + it;
            if (pix < prm.npix) {
                const T *fp = prm.feat + (size_t)pix * C + 4 * s;
                float f[M][4];
#pragma unroll
                for (int m = 0; m < M; ++m) load_elems<T, 4>(fp + 32 * m, f[m]);
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
#pragma unroll
                        for (int k = 0; k < 3; ++k) acc[it][k] = fmaf(f[m][e], wk[m][e][k], acc[it][k]);
            }
        }
        float z[3];
        butterfly8x3(acc, s, z);

        const uint32_t pix = p0 + lane;
        if (pix < prm.npix) {
            float x[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                x[k] = round_to<T>(sigmoidf_acc(z[k]));                  // bts_decoder.py:79 activation='sigmoid'
                store1(prm.coef_out + (size_t)pix * 3 + k, x[k]);
            }
            uint32_t row, j, b, i;
            prm.w.divmod(pix, row, j);
            prm.h.divmod(row, b, i);
            Angles a;
            decode_angles_for<T>(x[0], x[1], a);
            float n1[1] = {a.st * a.cp}, n2[1] = {a.st * a.sp}, n3[1] = {a.ct}, n4[1] = {x[2]};
            T *orow = prm.out + ((size_t)b * prm.out_sB + (size_t)(i * R) * prm.out_sH + j * R);
            T *drow = nullptr;
            if constexpr (D > 0) {
                if (prm.ds) drow = prm.ds + ((size_t)b * prm.ds_sB + (size_t)(i * NDS) * prm.ds_sH + j * NDS);
            }
            lpg_expand_store<T, R, 1, R, D, 0>(n1, n2, n3, n4, orow, prm.out_sH, drow, prm.ds_sH);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// TMA-staged forward.  The feature map is the only large input and 32 coarse pixels x C channels are
// ONE contiguous block of memory, so each warp streams its tiles through a private ring of NS
// shared-memory stages filled by 1-D bulk copies (cp.async.bulk, mbarrier completion): a chunk is
// 4 pixels x C channels (0.5-2 KB); lane 0 keeps NS-1 chunks in flight while the warp consumes the
// current one, so the HBM latency is hidden by the ring instead of by registers or occupancy.
// Lane (g, s) of a chunk reads pixel g, channels [32m+4s, 32m+4s+4): a quarter warp reads 128
// contiguous bytes per LDS (conflict-free).  After 8 chunks the 3-level butterfly leaves lane (g, s)
// with pixel 4s+g; one index shuffle puts pixel L on lane L for coalesced coefficient / depth stores.
// ------------------------------------------------------------------------------------------------
// A stage of a warp's ring is one bulk copy of 2-4 KB: CPB consecutive 4-pixel chunks.
// Ring geometry (measured: profiles/r01_sweep_head_ring.md, profiles/experiments/README.md):
//   forward             2 stages x 2 KB  (many resident warps: registers are the limit, a small ring keeps occupancy)
//   forward  r = 8 with 512-byte pixels (float32, C = 128)   3 stages x 4 KB: a lane expands 64 pixels -> ~100
//                       registers -> 16 warps per SM; with 2 x 2 KB those warps keep 2 KB each in flight, ~4 TB/s
//                       for the whole GPU (0.755 -> 0.840 of peak; narrower pixels measured slightly worse with it)
//   backward            3 stages x 4 KB  (it also writes g_feat, so it wants more bytes in flight per warp)
#ifndef BTSLPG_HEAD_FWD_STAGES
#define BTSLPG_HEAD_FWD_STAGES 2
#endif
#ifndef BTSLPG_HEAD_FWD_STAGE_BYTES
#define BTSLPG_HEAD_FWD_STAGE_BYTES 2048
#endif
#ifndef BTSLPG_HEAD_FWD8_STAGES
#define BTSLPG_HEAD_FWD8_STAGES 3
#endif
#ifndef BTSLPG_HEAD_FWD8_STAGE_BYTES
#define BTSLPG_HEAD_FWD8_STAGE_BYTES 4096
#endif
#ifndef BTSLPG_HEAD_BWD_STAGES
#define BTSLPG_HEAD_BWD_STAGES 3
#endif
#ifndef BTSLPG_HEAD_BWD_STAGE_BYTES
#define BTSLPG_HEAD_BWD_STAGE_BYTES 4096
#endif
template <typename T, int R, int M, bool FWD> struct HeadTmaCfg {
    static constexpr int kChunkPx = 4;                                        // pixels per chunk (= lane groups)
    static constexpr int kPxBytes = 32 * M * (int)sizeof(T);
    static constexpr bool kDeepFwd = FWD && R == 8 && kPxBytes >= 512;        // float32, C = 128
    static constexpr int kStages = !FWD ? BTSLPG_HEAD_BWD_STAGES : (kDeepFwd ? BTSLPG_HEAD_FWD8_STAGES : BTSLPG_HEAD_FWD_STAGES);
    static constexpr int kWantBytes = !FWD ? BTSLPG_HEAD_BWD_STAGE_BYTES : (kDeepFwd ? BTSLPG_HEAD_FWD8_STAGE_BYTES : BTSLPG_HEAD_FWD_STAGE_BYTES);
    static constexpr int kChunkBytes = kChunkPx * kPxBytes;
    static constexpr int kWant = kWantBytes / kChunkBytes;
    static constexpr int kCPB = kWant < 1 ? 1 : (kWant > 8 ? 8 : kWant);
    static constexpr int kStageBytes = kCPB * kChunkBytes;
    static constexpr int kStagesPerTile = 8 / kCPB;
};

template <typename T, int R, int M, bool FWD> __host__ __device__ constexpr int head_tma_smem_bytes(int warps) {
    return warps * HeadTmaCfg<T, R, M, FWD>::kStages * (HeadTmaCfg<T, R, M, FWD>::kStageBytes + 8);
}

// Per-warp ring of bulk-copied feature stages.  Stage sq (counted per warp) holds pixels
// [px0, px0 + 4*CPB) of tile sq / SPT; lane 0 issues, all lanes wait on the stage's mbarrier.
template <typename T, int R, int M, bool FWD> struct FeatRing {
    using Cfg = HeadTmaCfg<T, R, M, FWD>;
    static constexpr int C = 32 * M;
    unsigned char *ring;
    uint64_t *bars;
    const T *feat;
    uint32_t npix, warp, nwarps, nstages;

    __device__ __forceinline__ void init(unsigned char *smem, int wid, int nw, int lane, const T *feat_, uint32_t npix_, uint32_t warp_,
                                         uint32_t nwarps_, uint32_t ntiles) {
        ring = smem + (size_t)wid * Cfg::kStages * Cfg::kStageBytes;
        bars = reinterpret_cast<uint64_t *>(smem + (size_t)nw * Cfg::kStages * Cfg::kStageBytes) + wid * Cfg::kStages;
        feat = feat_; npix = npix_; warp = warp_; nwarps = nwarps_;
        nstages = ntiles * Cfg::kStagesPerTile;
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < Cfg::kStages; ++k) mbar_init(&bars[k], 1);
            mbar_fence_init();
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < Cfg::kStages; ++k) issue(k);
        }
    }
    __device__ __forceinline__ uint32_t stage_px0(uint32_t sq) const {
        return (warp + (sq / Cfg::kStagesPerTile) * nwarps) * 32 + (sq % Cfg::kStagesPerTile) * (Cfg::kCPB * Cfg::kChunkPx);
    }
    __device__ __forceinline__ void issue(uint32_t sq) {      // lane 0 only
        if (sq < nstages) {
            const uint32_t px0 = stage_px0(sq);
            if (px0 < npix) {
                const uint32_t npx = min((uint32_t)(Cfg::kCPB * Cfg::kChunkPx), npix - px0);
                uint64_t *bar = &bars[sq % Cfg::kStages];
                mbar_arrive_expect_tx(bar, npx * Cfg::kPxBytes);
                bulk_g2s(ring + (sq % Cfg::kStages) * Cfg::kStageBytes, feat + (size_t)px0 * C, npx * Cfg::kPxBytes, bar);
            }
        }
    }
    __device__ __forceinline__ void wait(uint32_t sq) { mbar_wait(&bars[sq % Cfg::kStages], (sq / Cfg::kStages) & 1); }
    // every lane is done with stage sq: refill it with stage sq + kStages
    __device__ __forceinline__ void release(uint32_t sq, int lane) {
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async();
            issue(sq + Cfg::kStages);
        }
    }
    // chunk c of stage sq, pixel g, channel offset ch
    __device__ __forceinline__ const unsigned char *at(uint32_t sq, int c, int g, int ch) const {
        return ring + (sq % Cfg::kStages) * Cfg::kStageBytes + c * Cfg::kChunkBytes + g * Cfg::kPxBytes + ch * (int)sizeof(T);
    }
};

template <typename T, int R, int D, int M>
__global__ void __launch_bounds__(256) head_lpg_fwd_tma_kernel(const __grid_constant__ HeadFwdParams<T> prm) {
    using Cfg = HeadTmaCfg<T, R, M, true>;
    constexpr int NDS = D ? R / D : 0;
    extern __shared__ __align__(128) unsigned char head_smem[];

    const int lane = threadIdx.x & 31, s = lane & 7, g = lane >> 3, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t ntiles = warp < prm.iters ? (prm.iters - warp + nwarps - 1) / nwarps : 0;
    FeatRing<T, R, M, true> fr;
    fr.init(head_smem, wid, nw, lane, prm.feat, prm.npix, warp, nwarps, ntiles);

    float wk[M][4][3];
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int k = 0; k < 3; ++k) wk[m][e][k] = __ldg(prm.kernel + (32 * m + 4 * s + e) * 3 + k);

    uint32_t sq = 0;
    for (uint32_t t = 0; t < ntiles; ++t) {
        const uint32_t p0 = (warp + t * nwarps) * 32;
        float acc[8][3];
#pragma unroll
        for (int st = 0; st < Cfg::kStagesPerTile; ++st, ++sq) {
            const uint32_t spx0 = p0 + st * (Cfg::kCPB * Cfg::kChunkPx);
            const bool live = spx0 < prm.npix;                      // warp-uniform
            if (live) fr.wait(sq);
#pragma unroll
            for (int c = 0; c < Cfg::kCPB; ++c) {
                const int it = st * Cfg::kCPB + c;
                acc[it][0] = acc[it][1] = acc[it][2] = 0.0f;
                if (live && spx0 + c * Cfg::kChunkPx + g < prm.npix) {
                    float f[M][4];
#pragma unroll
                    for (int m = 0; m < M; ++m) lds_elems<T, 4>(fr.at(sq, c, g, 32 * m + 4 * s), f[m]);
#pragma unroll
                    for (int m = 0; m < M; ++m)
#pragma unroll
                        for (int e = 0; e < 4; ++e)
#pragma unroll
                            for (int k = 0; k < 3; ++k) acc[it][k] = fmaf(f[m][e], wk[m][e][k], acc[it][k]);
                }
            }
            if (live) fr.release(sq, lane);
        }
        float z[3];
        butterfly8x3(acc, s, z);                                    // lane (g, s) holds pixel p0 + 4 s + g
        const int src_lane = 8 * (lane & 3) + (lane >> 2);          // lane L takes pixel p0 + L
#pragma unroll
        for (int k = 0; k < 3; ++k) z[k] = __shfl_sync(0xffffffffu, z[k], src_lane);

        const uint32_t pix = p0 + lane;
        if (pix < prm.npix) {
            float x[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                x[k] = round_to<T>(sigmoidf_acc(z[k]));                  // bts_decoder.py:79 activation='sigmoid'
                store1(prm.coef_out + (size_t)pix * 3 + k, x[k]);
            }
            uint32_t row, j, b, i;
            prm.w.divmod(pix, row, j);
            prm.h.divmod(row, b, i);
            Angles a;
            decode_angles_for<T>(x[0], x[1], a);
            float n1[1] = {a.st * a.cp}, n2[1] = {a.st * a.sp}, n3[1] = {a.ct}, n4[1] = {x[2]};
            T *orow = prm.out + ((size_t)b * prm.out_sB + (size_t)(i * R) * prm.out_sH + j * R);
            T *drow = nullptr;
            if constexpr (D > 0) {
                if (prm.ds) drow = prm.ds + ((size_t)b * prm.ds_sB + (size_t)(i * NDS) * prm.ds_sH + j * NDS);
            }
            lpg_expand_store<T, R, 1, R, D, 0>(n1, n2, n3, n4, orow, prm.out_sH, drow, prm.ds_sH);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// CTAs -> result for g_kernel: the last CTA to finish adds the per-CTA partial sums [nblk][C*3] in a
// fixed order.  The 256 threads are laid out as (C*3/4 float4 columns) x (S slices of the CTA range); a
// thread walks its slice with eight 16-byte L2 loads in flight (a one-entry-per-thread loop is a chain of
// ~300 dependent L2 round trips, several microseconds at the end of a 30-200 us kernel), then the S slice
// sums are combined in slice order through shared memory.  Order: CTAs s, s+S, s+2S, ... inside slice s,
// then slices 0..S-1 -- fixed for a given grid size, no atomics.
// ------------------------------------------------------------------------------------------------
template <int C> __device__ __forceinline__ void head_reduce_partials(const float *partial, uint32_t nblk, float *g_kernel) {
    constexpr int NCOL = C * 3 / 4;          // 24, 48 or 96 float4 columns
    constexpr int S = 256 / NCOL;            // 10, 5 or 2 slices
    __shared__ float4 comb[S][NCOL];
    const int col = threadIdx.x % NCOL, sl = threadIdx.x / NCOL;
    if (sl < S) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 *p4 = reinterpret_cast<const float4 *>(partial) + col;
        uint32_t blk = sl;
        for (; blk + 7 * S < nblk; blk += 8 * S) {
            float4 t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = __ldcg(p4 + (size_t)(blk + u * S) * NCOL);
#pragma unroll
            for (int u = 0; u < 8; ++u) { v.x += t[u].x; v.y += t[u].y; v.z += t[u].z; v.w += t[u].w; }
        }
        for (; blk < nblk; blk += S) {
            const float4 t = __ldcg(p4 + (size_t)blk * NCOL);
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        comb[sl][col] = v;
    }
    __syncthreads();
    if (threadIdx.x < NCOL) {
        float4 v = comb[0][threadIdx.x];
#pragma unroll
        for (int q = 1; q < S; ++q) {
            const float4 t = comb[q][threadIdx.x];
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        float *o = g_kernel + 4 * threadIdx.x;        // g_kernel may be a 4-byte-aligned slice of a gradient bucket
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
}

// ------------------------------------------------------------------------------------------------
// TMA-staged backward: same per-warp feature ring (feat is re-read for g_kernel); the patch gradients
// and saved coefficients of a tile are small and loaded directly by lane L for pixel L.
// ------------------------------------------------------------------------------------------------
template <typename T, int R, int D, int M>
__global__ void __launch_bounds__(256) head_lpg_bwd_tma_kernel(const __grid_constant__ HeadBwdParams<T> prm) {
    using Cfg = HeadTmaCfg<T, R, M, false>;
    constexpr int C = 32 * M;
    constexpr int NDS = D ? R / D : 0;
    constexpr int kMaxWarps = 8;
    extern __shared__ __align__(128) unsigned char head_smem[];
    __shared__ float red[kMaxWarps][C * 3];
    __shared__ bool is_last;

    const int lane = threadIdx.x & 31, s = lane & 7, g = lane >> 3, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const bool want_gk = prm.g_kernel != nullptr, want_gf = prm.g_feat != nullptr;
    const uint32_t ntiles = warp < prm.iters ? (prm.iters - warp + nwarps - 1) / nwarps : 0;
    FeatRing<T, R, M, false> fr;
    fr.init(head_smem, wid, nw, lane, prm.feat, prm.npix, warp, nwarps, want_gk ? ntiles : 0);

    float wk[M][4][3], dw[M][4][3];
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                wk[m][e][k] = want_gf ? __ldg(prm.kernel + (32 * m + 4 * s + e) * 3 + k) : 0.0f;
                dw[m][e][k] = 0.0f;
            }

    uint32_t sq = 0;
    for (uint32_t t = 0; t < ntiles; ++t) {
        const uint32_t p0 = (warp + t * nwarps) * 32;
        const uint32_t pix = p0 + lane;
        float dz[3] = {0.0f, 0.0f, 0.0f};
        if (pix < prm.npix) {                                       // lane L: LPG backward of pixel p0 + L
            uint32_t row, j, b, i;
            prm.w.divmod(pix, row, j);
            prm.h.divmod(row, b, i);
            float G[R][R];
            const T *grow = prm.g_full ? prm.g_full + ((size_t)b * prm.gf_sB + (size_t)(i * R) * prm.gf_sH + j * R) : nullptr;
            const T *drow = nullptr;
            if constexpr (D > 0) {
                if (prm.g_ds) drow = prm.g_ds + ((size_t)b * prm.gd_sB + (size_t)(i * NDS) * prm.gd_sH + j * NDS);
            }
            float x[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) x[k] = load1(prm.coef + (size_t)pix * 3 + k);
            lpg_load_patch<T, R, 1, R, D, 0>(grow, prm.gf_sH, drow, prm.gd_sH, G);
            float gc[3], acc[4];
            Angles a;
            decode_angles_for<T>(x[0], x[1], a);
            lpg_patch_partial<R, 1, R, 0>(G, 0, a.st * a.cp, a.st * a.sp, a.ct, acc);
            lpg_finish_grad(a, x[2], acc, gc);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (prm.g_coef_out) store1(prm.g_coef_out + (size_t)pix * 3 + k, gc[k]);
                dz[k] = gc[k] * x[k] * (1.0f - x[k]);                    // sigmoid'
            }
        }
#pragma unroll
        for (int st = 0; st < Cfg::kStagesPerTile; ++st) {
            const uint32_t spx0 = p0 + st * (Cfg::kCPB * Cfg::kChunkPx);
            const bool live = spx0 < prm.npix;                      // warp-uniform
            if (live && want_gk) fr.wait(sq);
#pragma unroll
            for (int c = 0; c < Cfg::kCPB; ++c) {
                const int it = st * Cfg::kCPB + c;
                float dzb[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) dzb[k] = __shfl_sync(0xffffffffu, dz[k], 4 * it + g);   // pixel p0 + 4 it + g lives on that lane
                const uint32_t pg = spx0 + c * Cfg::kChunkPx + g;
                if (live && pg < prm.npix) {
                    if (want_gk) {
                        float f[M][4];
#pragma unroll
                        for (int m = 0; m < M; ++m) lds_elems<T, 4>(fr.at(sq, c, g, 32 * m + 4 * s), f[m]);
#pragma unroll
                        for (int m = 0; m < M; ++m)
#pragma unroll
                            for (int e = 0; e < 4; ++e)
#pragma unroll
                                for (int k = 0; k < 3; ++k) dw[m][e][k] = fmaf(f[m][e], dzb[k], dw[m][e][k]);
                    }
                    if (want_gf) {
                        T *gp = prm.g_feat + (size_t)pg * C + 4 * s;
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            float gf[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                gf[e] = fmaf(dzb[2], wk[m][e][2], fmaf(dzb[1], wk[m][e][1], dzb[0] * wk[m][e][0]));
                            store_elems<T, 4>(gp + 32 * m, gf);
                        }
                    }
                }
            }
            if (want_gk) {
                if (live) fr.release(sq, lane);
                ++sq;
            }
        }
    }

    if (!want_gk) return;   // uniform across the grid

    // lanes -> warp: add the four lane groups (fixed order), lanes 0..7 then hold the warp totals
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float v = dw[m][e][k];
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (lane < 8) red[wid][(32 * m + 4 * s + e) * 3 + k] = v;
            }
    __syncthreads();
    for (int tt = threadIdx.x; tt < C * 3; tt += blockDim.x) {      // warps -> CTA
        float v = 0.0f;
        for (int w = 0; w < nw; ++w) v += red[w][tt];
        prm.partial[(size_t)blockIdx.x * (C * 3) + tt] = v;
    }
    __threadfence();                                                // CTAs -> result: the last CTA sums in CTA order
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(prm.counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        head_reduce_partials<C>(prm.partial, gridDim.x, prm.g_kernel);
        if (threadIdx.x == 0) *prm.counter = 0u;   // leave the workspace header zero for the next launch
    }
}

template <typename T, int R, int D, int M>
__global__ void __launch_bounds__(256) head_lpg_bwd_kernel(const __grid_constant__ HeadBwdParams<T> prm) {
    constexpr int C = 32 * M;
    constexpr int NDS = D ? R / D : 0;
    constexpr int kMaxWarps = 8;
    __shared__ float red[kMaxWarps][C * 3];
    __shared__ bool is_last;

    const int lane = threadIdx.x & 31, s = lane & 7, gbase = lane & 24, wid = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const bool want_gk = prm.g_kernel != nullptr, want_gf = prm.g_feat != nullptr;

    float wk[M][4][3], dw[M][4][3];
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                wk[m][e][k] = want_gf ? __ldg(prm.kernel + (32 * m + 4 * s + e) * 3 + k) : 0.0f;
                dw[m][e][k] = 0.0f;
            }

    for (uint32_t iter = warp; iter < prm.iters; iter += nwarps) {
        const uint32_t p0 = iter * 32;
        const uint32_t pix = p0 + lane;
        float dz[3] = {0.0f, 0.0f, 0.0f};
        if (pix < prm.npix) {
            uint32_t row, j, b, i;
            prm.w.divmod(pix, row, j);
            prm.h.divmod(row, b, i);
            float G[R][R];
            const T *grow = prm.g_full ? prm.g_full + ((size_t)b * prm.gf_sB + (size_t)(i * R) * prm.gf_sH + j * R) : nullptr;
            const T *drow = nullptr;
            if constexpr (D > 0) {
                if (prm.g_ds) drow = prm.g_ds + ((size_t)b * prm.gd_sB + (size_t)(i * NDS) * prm.gd_sH + j * NDS);
            }
            float x[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) x[k] = load1(prm.coef + (size_t)pix * 3 + k);
            lpg_load_patch<T, R, 1, R, D, 0>(grow, prm.gf_sH, drow, prm.gd_sH, G);
            float gc[3], acc[4];
            Angles a;
            decode_angles_for<T>(x[0], x[1], a);
            lpg_patch_partial<R, 1, R, 0>(G, 0, a.st * a.cp, a.st * a.sp, a.ct, acc);
            lpg_finish_grad(a, x[2], acc, gc);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (prm.g_coef_out) store1(prm.g_coef_out + (size_t)pix * 3 + k, gc[k]);
                dz[k] = gc[k] * x[k] * (1.0f - x[k]);                    // sigmoid'
            }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            float dzb[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) dzb[k] = __shfl_sync(0xffffffffu, dz[k], gbase + it);
            const uint32_t pg = p0 + gbase + it;
            if (pg < prm.npix) {
                if (want_gk) {
                    const T *fp = prm.feat + (size_t)pg * C + 4 * s;
                    float f[M][4];
#pragma unroll
                    for (int m = 0; m < M; ++m) load_elems<T, 4>(fp + 32 * m, f[m]);
#pragma unroll
                    for (int m = 0; m < M; ++m)
#pragma unroll
                        for (int e = 0; e < 4; ++e)
#pragma unroll
                            for (int k = 0; k < 3; ++k) dw[m][e][k] = fmaf(f[m][e], dzb[k], dw[m][e][k]);
                }
                if (want_gf) {
                    T *gp = prm.g_feat + (size_t)pg * C + 4 * s;
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        float gf[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            gf[e] = fmaf(dzb[2], wk[m][e][2], fmaf(dzb[1], wk[m][e][1], dzb[0] * wk[m][e][0]));
                        store_elems<T, 4>(gp + 32 * m, gf);
                    }
                }
            }
        }
    }

    if (!want_gk) return;   // uniform across the grid

    // lanes -> warp: add the four lane groups (fixed order), lanes 0..7 then hold the warp totals
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float v = dw[m][e][k];
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (lane < 8) red[wid][(32 * m + 4 * s + e) * 3 + k] = v;
            }
    __syncthreads();
    // warps -> CTA
    const int nw = blockDim.x >> 5;
    for (int t = threadIdx.x; t < C * 3; t += blockDim.x) {
        float v = 0.0f;
        for (int w = 0; w < nw; ++w) v += red[w][t];
        prm.partial[(size_t)blockIdx.x * (C * 3) + t] = v;
    }
    // CTAs -> result: the last CTA to arrive sums all partials in CTA order
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(prm.counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        head_reduce_partials<C>(prm.partial, gridDim.x, prm.g_kernel);
        if (threadIdx.x == 0) *prm.counter = 0u;   // leave the workspace header zero for the next launch
    }
}

// ------------------------------------------------------------------------------------------------
// Generic head kernels: any channel count / strides.  One thread per pixel; used when C is not
// 32, 64 or 128 or when feat is not contiguous.  Deterministic (serial sums, fixed-order tree).
// ------------------------------------------------------------------------------------------------
template <typename T> struct HeadGenericParams {
    const T *feat;
    int64_t f_sP, f_sC;   // element strides: pixel (contiguous pixels required), channel
    const float *kernel;
    const T *coef;        // bwd: saved sigmoid output (npix,3) contiguous
    T *coef_out;          // fwd
    const T *g_coef;      // bwd: LPG coefficient gradient (npix,3) contiguous
    T *g_feat;
    int64_t gf_sP, gf_sC;
    float *g_kernel;
    int64_t npix;
    int32_t C;
};

template <typename T> __global__ void __launch_bounds__(128) head_fwd_generic_kernel(const __grid_constant__ HeadGenericParams<T> prm) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= prm.npix) return;
    float z[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < prm.C; ++c) {
        const float f = load1(prm.feat + p * prm.f_sP + c * prm.f_sC);
#pragma unroll
        for (int k = 0; k < 3; ++k) z[k] = fmaf(f, __ldg(prm.kernel + c * 3 + k), z[k]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) store1(prm.coef_out + p * 3 + k, sigmoidf_acc(z[k]));
}

template <typename T> __global__ void __launch_bounds__(128) head_bwd_feat_generic_kernel(const __grid_constant__ HeadGenericParams<T> prm) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= prm.npix) return;
    float dz[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float x = load1(prm.coef + p * 3 + k);
        dz[k] = load1(prm.g_coef + p * 3 + k) * x * (1.0f - x);
    }
    for (int c = 0; c < prm.C; ++c) {
        const float v = fmaf(dz[2], __ldg(prm.kernel + c * 3 + 2), fmaf(dz[1], __ldg(prm.kernel + c * 3 + 1), dz[0] * __ldg(prm.kernel + c * 3)));
        store1(prm.g_feat + p * prm.gf_sP + c * prm.gf_sC, v);
    }
}

// one CTA per channel; threads stride over pixels, then a fixed-order shared-memory tree
template <typename T> __global__ void __launch_bounds__(256) head_bwd_kernel_generic_kernel(const __grid_constant__ HeadGenericParams<T> prm) {
    __shared__ float red[3][256];
    const int c = blockIdx.x;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int64_t p = threadIdx.x; p < prm.npix; p += blockDim.x) {
        const float f = load1(prm.feat + p * prm.f_sP + c * prm.f_sC);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float x = load1(prm.coef + p * 3 + k);
            acc[k] = fmaf(f, load1(prm.g_coef + p * 3 + k) * x * (1.0f - x), acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) red[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int stride = 128; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride)
#pragma unroll
            for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + stride];
        __syncthreads();
    }
    if (threadIdx.x < 3) prm.g_kernel[c * 3 + threadIdx.x] = red[threadIdx.x][0];
}

}  // namespace btslpg
