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
// CTAs -> result in TWO levels.  With one level the last CTA walks all ~300 partial rows: each of its threads owns a column
// and can keep only ~8 loads in flight, a chain of ~20-40 dependent L2 round trips -- 5-8 us at the end of a 30-45 us kernel
// (the fixed cost that held the r = 8 backward at 0.55-0.66 of the HBM peak while r = 4 / r = 2 reached 0.85-0.88).
// Here CTAs form groups of kHeadGroup consecutive block indices with one completion counter each: the last CTA of a group
// to finish adds the group's rows (16 independent loads per column: one round trip) into a group row, and the last group to
// finish adds the <= 64 group rows.  Every sum has a fixed order (rows of a group in block order, groups in order), only
// integer counters are atomic, so g_kernel stays bit-reproducible for a given grid size.
// Workspace: counter[0] = finished groups, counter[1 + g] = finished CTAs of group g (all left zero on exit);
// partial rows [0, nblk) = CTAs, [nblk, nblk + ngroups) = groups.
// Call with the CTA's own row already written by its threads (any thread layout); contains CTA barriers.
// ------------------------------------------------------------------------------------------------
constexpr int kHeadGroup = 16;
constexpr int kHeadMaxGroups = 62;          // counters live in the 256-byte workspace header

template <int C> __device__ __forceinline__ void head_grid_reduce(float *partial, unsigned int *counter, float *g_kernel) {
    constexpr int N = C * 3;
    __shared__ int s_flag;
    const uint32_t nblk = gridDim.x, ngroups = (nblk + kHeadGroup - 1) / kHeadGroup;
    const uint32_t grp = blockIdx.x / kHeadGroup;
    const uint32_t first = grp * kHeadGroup, count = min((uint32_t)kHeadGroup, nblk - first);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_flag = (atomicAdd(counter + 1 + grp, 1u) == count - 1);
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    for (int c = threadIdx.x; c < N; c += blockDim.x) {             // rows of the group, block order
        float v[kHeadGroup];
#pragma unroll
        for (int r = 0; r < kHeadGroup; ++r) v[r] = (uint32_t)r < count ? __ldcg(partial + (size_t)(first + r) * N + c) : 0.0f;
        float sum = v[0];
#pragma unroll
        for (int r = 1; r < kHeadGroup; ++r) sum += v[r];
        partial[(size_t)(nblk + grp) * N + c] = sum;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        counter[1 + grp] = 0u;                                      // nobody else touches this group's counter any more
        s_flag = (atomicAdd(counter, 1u) == ngroups - 1);
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    for (int c = threadIdx.x; c < N; c += blockDim.x) {             // group rows, group order
        float sum = 0.0f;
        uint32_t g0 = 0;
        for (; g0 + 16 <= ngroups; g0 += 16) {
            float v[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = __ldcg(partial + (size_t)(nblk + g0 + r) * N + c);
#pragma unroll
            for (int r = 0; r < 16; ++r) sum += v[r];
        }
        for (; g0 < ngroups; ++g0) sum += __ldcg(partial + (size_t)(nblk + g0) * N + c);
        g_kernel[c] = sum;                                          // g_kernel may be a 4-byte-aligned slice of a gradient bucket
    }
    if (threadIdx.x == 0) *counter = 0u;                            // leave the workspace header zero for the next launch
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
    head_grid_reduce<C>(prm.partial, prm.counter, prm.g_kernel);    // CTAs -> groups -> result, fixed order
}

// ------------------------------------------------------------------------------------------------
// r = 8 backward, lane-split form.  The kernel above gives a lane the whole 8 x 8 patch of its coarse pixel (64 gradient
// registers next to 2 x 12*M weight / accumulator registers) and hands a warp 32 pixels at a time: 153 600 pixels at
// B = 32, 480x640 are 4800 warp-tiles for ~1200 resident warps -- 4.05 rounds, i.e. one fifth of the machine idles in the
// last round -- and the 64-register patch load sits un-prefetched in front of every tile (0.55-0.66 of the HBM peak).
// Here a tile is 8 coarse pixels (19 200 tiles: 16.2 rounds, 5 % tail):
//   phase A  lane (q = lane / 4, sub = lane % 4) owns patch rows 2 sub, 2 sub + 1 of pixel q: 16 gradient registers,
//            two 32-byte loads whose NEXT tile's instance is issued before this tile's arithmetic (software prefetch, 21
//            registers); the four partial sums of a pixel are added by a fixed xor tree over lane bits 0, 1 -- every
//            lane of the quad ends with the same bits, (s0 + s1) + (s2 + s3): deterministic, no atomics.
//            The lane's two table rows are a per-LANE index, which the constant bank cannot serve without serialising;
//            they come from a 288-byte shared-memory copy of the table (w rows and row offsets; u = a*w and v = b*w are
//            recomputed with the multiplication that generated the table, bit for bit).
//   phase B  as above (8 lanes per pixel over the channels, 4 pixels per step, features from the per-warp TMA ring whose
//            stage is now exactly one tile); the head kernel for g_feat is read from shared memory instead of 12*M registers.
// ------------------------------------------------------------------------------------------------
#ifndef BTSLPG_HEAD_BWD8_RING_BYTES
#define BTSLPG_HEAD_BWD8_RING_BYTES 12288      // bytes of feature stages a warp keeps in flight
#endif
template <typename T, int M> struct Bwd8Cfg {
    static constexpr int kTilePx = 8;
    static constexpr int kPxBytes = 32 * M * (int)sizeof(T);
    static constexpr int kStageBytes = kTilePx * kPxBytes;                 // 4 KB (float32, C = 128) ... 512 B (bfloat16, C = 32)
    static constexpr int kWantStages = BTSLPG_HEAD_BWD8_RING_BYTES / kStageBytes;
    static constexpr int kStages = kWantStages < 3 ? 3 : (kWantStages > 8 ? 8 : kWantStages);
};
template <typename T, int M> __host__ __device__ constexpr int head_bwd8_smem_bytes(int warps) {
    return warps * Bwd8Cfg<T, M>::kStages * (Bwd8Cfg<T, M>::kStageBytes + 8);
}

// What lane (q, sub) needs of a tile, as RAW words: nothing may depend on a load before the tile that uses it (a widening
// shift, or adding the down-sampled gradient in, right behind the load would stall the warp for a DRAM round trip per tile --
// ncu on the first version: long_scoreboard 3.9 of 5.6 stall cycles per issue, DRAM 45 % busy).
template <typename T, int D> struct Bwd8Patch {
    static constexpr int GW = 8 * (int)sizeof(T) / 4;       // words per 8-pixel patch row: 8 (float32) / 4 (bfloat16)
    uint32_t g[2][GW];
    uint32_t ds[2];                                         // float32: 2 words; bfloat16: one word holding both
    uint32_t x[3];                                          // coefficient: float32 bits / zero-extended bfloat16
    bool valid, has_ds;
};

__device__ __forceinline__ uint32_t ldg_u16(const void *p) {
    unsigned short v;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return (uint32_t)v;
}

// loads of lane (q, sub) for the tile that starts at pixel p0 (issued one tile ahead of their use)
template <typename T, int D>
__device__ __forceinline__ void bwd8_load(const HeadBwdParams<T> &prm, uint32_t p0, int q, int sub, Bwd8Patch<T, D> &P) {
    constexpr int GW = Bwd8Patch<T, D>::GW;
    const uint32_t pix = p0 + q;
    P.valid = pix < prm.npix;
    P.has_ds = false;
    if (!P.valid) return;
    uint32_t row, j, b, i;
    prm.w.divmod(pix, row, j);
    prm.h.divmod(row, b, i);
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int k = 0; k < 3; ++k) ldg_nc_l1<1>(prm.coef + (size_t)pix * 3 + k, &P.x[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) P.x[k] = ldg_u16(prm.coef + (size_t)pix * 3 + k);
    }
    if (prm.g_full) {
        const T *grow = prm.g_full + ((size_t)b * prm.gf_sB + (size_t)(i * 8 + 2 * sub) * prm.gf_sH + j * 8);
        ldg_nc<GW>(grow, P.g[0]);
        ldg_nc<GW>(grow + prm.gf_sH, P.g[1]);
    } else {
#pragma unroll
        for (int e = 0; e < GW; ++e) P.g[0][e] = P.g[1][e] = 0u;
    }
    if constexpr (D > 0) {
        // bts_decoder.py:81 scatter: the down-sampled gradient lands on patch rows 0 and 4 (sub 0 and 2, first row), columns 0 and 4
        if (prm.g_ds && (sub & 1) == 0) {
            P.has_ds = true;
            const T *dp = prm.g_ds + ((size_t)b * prm.gd_sB + (size_t)(i * 2 + (sub >> 1)) * prm.gd_sH + j * 2);
            if constexpr (sizeof(T) == 4) ldg_nc<2>(dp, P.ds);
            else ldg_nc<1>(dp, P.ds);
        }
    }
}

// raw words -> the lane's two rows of gradients (down-sampled gradient added in) and its coefficient
template <typename T, int D>
__device__ __forceinline__ void bwd8_unpack(const Bwd8Patch<T, D> &P, float (&G)[2][8], float (&x)[3]) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int e = 0; e < 8; ++e) G[k][e] = __uint_as_float(P.g[k][e]);
#pragma unroll
        for (int k = 0; k < 3; ++k) x[k] = __uint_as_float(P.x[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                G[k][2 * e] = bf16_lo(P.g[k][e]);
                G[k][2 * e + 1] = bf16_hi(P.g[k][e]);
            }
#pragma unroll
        for (int k = 0; k < 3; ++k) x[k] = __uint_as_float(P.x[k] << 16);
    }
    if constexpr (D > 0) {
        if (P.has_ds) {
            if constexpr (sizeof(T) == 4) {
                G[0][0] += __uint_as_float(P.ds[0]);
                G[0][4] += __uint_as_float(P.ds[1]);
            } else {
                G[0][0] += bf16_lo(P.ds[0]);
                G[0][4] += bf16_hi(P.ds[0]);
            }
        }
    }
}

template <typename T, int D, int M>
__global__ void __launch_bounds__(256, 2) head_lpg_bwd8_kernel(const __grid_constant__ HeadBwdParams<T> prm) {
    using Cfg = Bwd8Cfg<T, M>;
    using Tab = DirTable<8>;
    constexpr int C = 32 * M;
    constexpr int kMaxWarps = 8;
    extern __shared__ __align__(128) unsigned char head_smem[];
    __shared__ __align__(16) float s_w[64];          // direction table rows (w), indexed by the lane's rows
    __shared__ float s_off[8];
    __shared__ __align__(16) float s_wk[C * 3];      // head kernel [C][3] for g_feat

    const int lane = threadIdx.x & 31, s = lane & 7, g = lane >> 3, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int q = lane >> 2, sub = lane & 3;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const bool want_gk = prm.g_kernel != nullptr, want_gf = prm.g_feat != nullptr;
    const uint32_t tiles_total = (prm.npix + Cfg::kTilePx - 1) / Cfg::kTilePx;
    const uint32_t ntiles = warp < tiles_total ? (tiles_total - warp + nwarps - 1) / nwarps : 0;

    // per-warp ring of one-tile feature stages.  The kernel is short (30-45 us at B = 32), so the prologue is ordered by
    // latency: every warp starts its own feature copies and its first patch loads BEFORE the CTA-wide table set-up, whose
    // global loads (the head kernel) would otherwise put one more DRAM round trip in front of the first copy.
    unsigned char *ring = head_smem + (size_t)wid * Cfg::kStages * Cfg::kStageBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(head_smem + (size_t)nw * Cfg::kStages * Cfg::kStageBytes) + wid * Cfg::kStages;
    const uint32_t nstages = want_gk ? ntiles : 0;
    auto issue = [&](uint32_t sq) {                  // lane 0 only
        if (sq < nstages) {
            const uint32_t px0 = (warp + sq * nwarps) * Cfg::kTilePx;
            const uint32_t npx = min((uint32_t)Cfg::kTilePx, prm.npix - px0);
            uint64_t *bar = &bars[sq % Cfg::kStages];
            mbar_arrive_expect_tx(bar, npx * Cfg::kPxBytes);
            bulk_g2s(ring + (sq % Cfg::kStages) * Cfg::kStageBytes, prm.feat + (size_t)px0 * C, npx * Cfg::kPxBytes, bar);
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < Cfg::kStages; ++k) mbar_init(&bars[k], 1);
        mbar_fence_init();
#pragma unroll
        for (int k = 0; k < Cfg::kStages; ++k) issue(k);
    }
    Bwd8Patch<T, D> cur;
    cur.valid = cur.has_ds = false;
    if (ntiles > 0) bwd8_load<T, D>(prm, warp * Cfg::kTilePx, q, sub, cur);

    if (threadIdx.x < 64) s_w[threadIdx.x] = c_w8[threadIdx.x];
    if (threadIdx.x < 8) s_off[threadIdx.x] = c_off8[threadIdx.x];
    for (int k = threadIdx.x; k < C * 3; k += blockDim.x) s_wk[k] = __ldg(prm.kernel + k);
    __syncthreads();                                 // tables / kernel copy visible to the CTA; lane 0's barrier set-up to its warp

    float dw[M][4][3];
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int k = 0; k < 3; ++k) dw[m][e][k] = 0.0f;

    // one tile: phase A from the prefetched words P, phase B from ring stage t
    auto process = [&](const Bwd8Patch<T, D> &P, uint32_t t) {
        const uint32_t p0 = (warp + t * nwarps) * Cfg::kTilePx;
        // ---- phase A: LPG backward of pixel p0 + q, rows 2 sub and 2 sub + 1 on this lane
        float dz[3] = {0.0f, 0.0f, 0.0f};
        {
            float x[3] = {0.5f, 0.5f, 0.0f}, G[2][8];
            if (P.valid) {
                bwd8_unpack<T, D>(P, G, x);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) G[0][e] = G[1][e] = 0.0f;
            }
            Angles a;
            decode_angles_for<T>(x[0], x[1], a);
            const float n1 = a.st * a.cp, n2 = a.st * a.sp, n3 = a.ct;
            const F2 n2b = f2(n2);
            F2 r1, r2, r3, r4;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float ak = s_off[2 * sub + k];
                const F2 A2 = f2(fmaf(ak, n1, n3)), ak2 = f2(ak);
                const float4 wa = *reinterpret_cast<const float4 *>(&s_w[(2 * sub + k) * 8]);
                const float4 wb = *reinterpret_cast<const float4 *>(&s_w[(2 * sub + k) * 8 + 4]);
                const F2 wq[4] = {f2(wa.x, wa.y), f2(wa.z, wa.w), f2(wb.x, wb.y), f2(wb.z, wb.w)};
#pragma unroll
                for (int c2 = 0; c2 < 4; ++c2) {
                    const F2 w = wq[c2];
                    const F2 inv = rcp2(fma2(w, fma2(Tab::off2(2 * c2), n2b, A2), f2(BTSLPG_EPS_F)));
                    const F2 tt = mul2(f2(G[k][2 * c2], G[k][2 * c2 + 1]), inv);
                    const F2 y = mul2(tt, inv);
                    const F2 v = mul2(Tab::off2(2 * c2), w);            // == the table's v = b_q * w
                    const F2 u = mul2(ak2, w);                          // == the table's u = a_p * w
                    if (k == 0 && c2 == 0) {
                        r4 = tt; r3 = mul2(y, w); r2 = mul2(y, v); r1 = mul2(y, u);
                    } else {
                        r4 = add2(r4, tt); r3 = fma2(y, w, r3); r2 = fma2(y, v, r2); r1 = fma2(y, u, r1);
                    }
                }
            }
            float acc[4] = {lo(r1) + hi(r1), lo(r2) + hi(r2), lo(r3) + hi(r3), lo(r4) + hi(r4)};
#pragma unroll
            for (int k = 0; k < 4; ++k) {                               // fixed tree over the quad: (s0 + s1) + (s2 + s3)
                acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1);
                acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2);
            }
            float gc[3];
            lpg_finish_grad(a, x[2], acc, gc);
            if (P.valid) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (prm.g_coef_out && sub == 0) store1(prm.g_coef_out + (size_t)(p0 + q) * 3 + k, gc[k]);
                    dz[k] = gc[k] * x[k] * (1.0f - x[k]);                // sigmoid'
                }
            }
        }

        // ---- phase B: g_feat and the g_kernel partial sums of the tile's 8 pixels, 4 pixels per step
        if (want_gk) mbar_wait(&bars[t % Cfg::kStages], (t / Cfg::kStages) & 1);
        const unsigned char *stage = ring + (t % Cfg::kStages) * Cfg::kStageBytes;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float dzb[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) dzb[k] = __shfl_sync(0xffffffffu, dz[k], 4 * (4 * c + g));    // pixel 4c + g lives on lanes 4(4c+g)..+3
            const uint32_t pg = p0 + 4 * c + g;
            if (pg < prm.npix) {
                if (want_gk) {
                    float f[M][4];
#pragma unroll
                    for (int m = 0; m < M; ++m) lds_elems<T, 4>(stage + (4 * c + g) * Cfg::kPxBytes + (32 * m + 4 * s) * (int)sizeof(T), f[m]);
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
                        const float4 *wp = reinterpret_cast<const float4 *>(&s_wk[(32 * m + 4 * s) * 3]);
                        const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];           // 4 channels x 3 taps, [c][k]
                        float gf[4];
                        gf[0] = fmaf(dzb[2], w0.z, fmaf(dzb[1], w0.y, dzb[0] * w0.x));
                        gf[1] = fmaf(dzb[2], w1.y, fmaf(dzb[1], w1.x, dzb[0] * w0.w));
                        gf[2] = fmaf(dzb[2], w2.x, fmaf(dzb[1], w1.w, dzb[0] * w1.z));
                        gf[3] = fmaf(dzb[2], w2.w, fmaf(dzb[1], w2.z, dzb[0] * w2.y));
                        store_elems<T, 4>(gp + 32 * m, gf);
                    }
                }
            }
        }
        if (want_gk) {                                   // every lane is done with the stage: refill it
            __syncwarp();
            if (lane == 0) {
                fence_proxy_async();
                issue(t + Cfg::kStages);
            }
        }
    };

    // The next tile's words are requested before this tile's arithmetic.  (Two alternating buffers without the register copy,
    // and packed FFMA2 accumulators for phase B, were measured: 51.7 us against 44.9 us for this form at C = 128 float32 --
    // the second prefetch buffer and the pair-aligned accumulators cost more under the 128-register cap than they save.)
    for (uint32_t t = 0; t < ntiles; ++t) {
        Bwd8Patch<T, D> nxt;
        nxt.valid = nxt.has_ds = false;
        if (t + 1 < ntiles) bwd8_load<T, D>(prm, (warp + (t + 1) * nwarps) * Cfg::kTilePx, q, sub, nxt);    // in flight during this tile
        process(cur, t);
        cur = nxt;
    }

    if (!want_gk) return;   // uniform across the grid

    // lanes -> warp -> CTA -> grid, as in the kernel above; the per-warp rows reuse the (now idle) ring
    __syncthreads();
    float (*red)[C * 3] = reinterpret_cast<float (*)[C * 3]>(head_smem);
    static_assert(kMaxWarps * C * 3 * 4 <= 8 * 3 * 512, "the reduction rows must fit the smallest ring (8 warps x 3 stages x 512 B)");
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
    head_grid_reduce<C>(prm.partial, prm.counter, prm.g_kernel);    // CTAs -> groups -> result, fixed order
}

template <typename T, int R, int D, int M>
__global__ void __launch_bounds__(256) head_lpg_bwd_kernel(const __grid_constant__ HeadBwdParams<T> prm) {
    constexpr int C = 32 * M;
    constexpr int NDS = D ? R / D : 0;
    constexpr int kMaxWarps = 8;
    __shared__ float red[kMaxWarps][C * 3];

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
    head_grid_reduce<C>(prm.partial, prm.counter, prm.g_kernel);    // CTAs -> groups -> result, fixed order
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
