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
#include "tail_kernels.cuh"   // tail_sigmoid
#include "tma_pipe.cuh"       // smem_u32

namespace btslpg {

__device__ __forceinline__ float load_smem1(const float *p) { return *p; }
__device__ __forceinline__ float load_smem1(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

constexpr int kDcThreads = 256;
constexpr int kDcTileW = 128;            // pixels of one image row per work item (C = 32)
// 16 channels: twice the pixels, i.e. the same 16 KB of x per stage (the per-item barriers and the strip set-up weigh twice as much on
// half the bytes otherwise)
template <int C> __host__ __device__ constexpr int dc_tile_w() { return C <= 16 ? 2 * kDcTileW : kDcTileW; }
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
    int vec_g;           // float32 g rows start on 16-byte boundaries (W % 4 == 0): 16-byte strip copies
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

// Two stages of (x tile, 3-row strip of g) in shared memory, filled with cp.async one work item ahead of the arithmetic.
// The first version loaded the g strip, synchronised, loaded x into registers and only then computed: two exposed DRAM round
// trips per 128-pixel item and CTA (0.45 of the HBM peak at C = 32, 0.34 at C = 16).  With the stages: 0.72 / 0.54; with 256-pixel tiles
// at C = 16, packed FMAs and the strip copied as 16-byte pieces (the 4-byte copies with per-element bounds logic were a quarter of the
// kernel's instructions): 0.78 / 0.71 (profiles/r02_tail_f32.json).
template <typename T, int C> struct DcBwdStage {
    static constexpr int kTW = dc_tile_w<C>();
    static constexpr int kXBytes = kTW * C * (int)sizeof(T);                  // 16 KB (float32)
    // a strip row: left halo pixel at index 3, the tile's own kTW pixels at 4 .. kTW + 3 (16-byte aligned: float32 rows whose width
    // is a multiple of 4 are copied as 16-byte pieces), right halo pixel at kTW + 4
    static constexpr int kGRow = kTW + 8;
    static constexpr int kGFloats = 3 * kGRow;
    static constexpr int kGBytes = (kGFloats * (int)sizeof(T) + 15) / 16 * 16;
    static constexpr int kBytes = kXBytes + kGBytes;
};
template <typename T, int C> __host__ __device__ constexpr int depthconv_bwd_smem_bytes() { return 2 * DcBwdStage<T, C>::kBytes; }

template <typename T, int C, bool ELU> __global__ void __launch_bounds__(kDcThreads, 2) depthconv_bwd_kernel(const __grid_constant__ DepthConvBwdParams<T> prm) {
    using St = DcBwdStage<T, C>;
    constexpr int LPP = C / 4;                 // lanes per pixel
    constexpr int PPW = 32 / LPP;              // pixels per warp pass
    constexpr int NW = kDcThreads / 32;
    constexpr int NCOL4 = 9 * C / 4;           // float4 columns of g_w: 72 (C = 32) or 36 (C = 16)
    constexpr int EPV = 16 / (int)sizeof(T);   // elements per 16-byte copy
    extern __shared__ __align__(16) unsigned char dcb_smem[];
    __shared__ __align__(16) float red[NW][9 * C];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int cg = lane % LPP, pl = lane / LPP;

    // the 36 weights of the thread's four channels and its 36 partial sums of g_w as packed float32 pairs: the 72 FMAs per pixel
    // issue as 36 FFMA2 (the kernel is bound by instruction issue: ncu 65 % of the issue slots at 16 warps per SM)
    F2 wr[9][2], acc[9][2];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            wr[t][h] = f2(__ldg(prm.w + t * C + 4 * cg + 2 * h), __ldg(prm.w + t * C + 4 * cg + 2 * h + 1));
            acc[t][h] = f2(0.0f);
        }

    // copies of one work item into a stage: the x tile (16-byte pieces) and the 3-row strip of g with its one-pixel halo
    // (element-wise, zero-filled outside the image: padding='same')
    auto prefetch = [&](uint32_t item, int stage) {
        uint32_t rowi, xb, b, y;
        prm.div_cb.divmod(item, rowi, xb);     // rowi = b * H + y
        prm.div_h.divmod(rowi, b, y);
        const uint32_t x0 = xb * St::kTW;
        const uint32_t npx = min((uint32_t)St::kTW, prm.W - x0);
        const uint32_t sbase = smem_u32(dcb_smem) + stage * St::kBytes;
        const T *xsrc = prm.x + ((size_t)rowi * prm.W + x0) * C;
        for (uint32_t i = threadIdx.x; i < npx * C / EPV; i += kDcThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + i * 16), "l"(xsrc + (size_t)i * EPV) : "memory");
        if (sizeof(T) == 4 && prm.vec_g) {
            // float32, W a multiple of 4: one 16-byte copy per thread for the three rows' own pixels (the source size trims the piece
            // at the right image border, 0 bytes = all zeros), six 4-byte copies for the halo pixels
            constexpr int NV = St::kTW / 4;
            static_assert(3 * NV <= kDcThreads - 6, "one strip copy per thread");
            if (threadIdx.x < 3 * NV) {
                const int r = threadIdx.x / NV, v = threadIdx.x % NV;
                const int yy = (int)y + r - 1, xx = (int)x0 + 4 * v;
                uint32_t nbytes = 0;
                const T *gsrc = prm.g;
                if (yy >= 0 && yy < (int)prm.H && xx < (int)prm.W) {
                    nbytes = 4u * (uint32_t)min(4, (int)prm.W - xx);
                    gsrc = prm.g + ((size_t)b * prm.H + yy) * prm.W + xx;
                }
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + St::kXBytes + (r * St::kGRow + 4 + 4 * v) * 4), "l"(gsrc),
                             "r"(nbytes) : "memory");
            } else if (threadIdx.x >= kDcThreads - 6) {
                const int hidx = threadIdx.x - (kDcThreads - 6), r = hidx >> 1, right = hidx & 1;
                const int yy = (int)y + r - 1, xx = right ? (int)x0 + St::kTW : (int)x0 - 1;
                const bool in = yy >= 0 && yy < (int)prm.H && xx >= 0 && xx < (int)prm.W;
                const T *gsrc = in ? prm.g + ((size_t)b * prm.H + yy) * prm.W + xx : prm.g;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sbase + St::kXBytes + (r * St::kGRow + (right ? St::kTW + 4 : 3)) * 4),
                             "l"(gsrc), "r"(in ? 4 : 0) : "memory");
            }
            return;
        }
        for (uint32_t i = threadIdx.x; i < 3 * (St::kTW + 2); i += kDcThreads) {
            const uint32_t r = i / (St::kTW + 2), j = i % (St::kTW + 2);
            const int yy = (int)y + (int)r - 1, xx = (int)x0 + (int)j - 1;
            const bool in = yy >= 0 && yy < (int)prm.H && xx >= 0 && xx < (int)prm.W && j <= npx + 1;
            const T *gsrc = in ? prm.g + ((size_t)b * prm.H + yy) * prm.W + xx : prm.g;
            const uint32_t slot = r * St::kGRow + j + 3;
            if constexpr (sizeof(T) == 4) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sbase + St::kXBytes + slot * 4), "l"(gsrc), "r"(in ? 4 : 0) : "memory");
            } else {                            // 2-byte elements: below cp.async's granularity, a plain store of the loaded value
                reinterpret_cast<T *>(dcb_smem + stage * St::kBytes + St::kXBytes)[slot] = in ? *gsrc : T(0.0f);
            }
        }
    };

    uint32_t item = blockIdx.x;
    if (item < prm.items) prefetch(item, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int k = 0; item < prm.items; item += gridDim.x, ++k) {
        const int stage = k & 1;
        if (item + gridDim.x < prm.items) prefetch(item + gridDim.x, stage ^ 1);       // the next item flies during this one's arithmetic
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        uint32_t rowi, xb;
        prm.div_cb.divmod(item, rowi, xb);
        const uint32_t x0 = xb * St::kTW;
        const uint32_t npx = min((uint32_t)St::kTW, prm.W - x0);
        const T *xs = reinterpret_cast<const T *>(dcb_smem + stage * St::kBytes);
        const T *gs = reinterpret_cast<const T *>(dcb_smem + stage * St::kBytes + St::kXBytes);       // [3][St::kGRow]
        const size_t row0 = ((size_t)rowi * prm.W + x0);
        for (uint32_t px = wid * PPW + pl; px < npx; px += NW * PPW) {
            float xv[4];
            lds_elems<T, 4>(xs + (size_t)px * C + 4 * cg, xv);
            F2 out2[2] = {f2(0.0f), f2(0.0f)};
            float dact[4];
            if constexpr (ELU) {
                // x is the PRE-activation of iconv1 (bts_decoder.py:100): the layer input is elu(x), d elu / d x = x > 0 ? 1 : exp(x)
                // (the same 4-instruction ELU as the forward, so that forward and backward see one function)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float ex = ex2_sfu(xv[e] * kLog2e);
                    dact[e] = xv[e] > 0.0f ? 1.0f : ex;
                    xv[e] = xv[e] > 0.0f ? xv[e] : ex - 1.0f;
                }
            }
            const F2 xv2[2] = {f2(xv[0], xv[1]), f2(xv[2], xv[3])};
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int tp = (dy + 1) * 3 + (dx + 1);
                    const F2 gv = f2(load_smem1(gs + (1 - dy) * St::kGRow + px + 4 - dx));      // g at pixel q - (dy, dx)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        out2[h] = fma2(gv, wr[tp][h], out2[h]);
                        acc[tp][h] = fma2(gv, xv2[h], acc[tp][h]);
                    }
                }
            float out[4];
            unpack(out2[0], out[0], out[1]);
            unpack(out2[1], out[2], out[3]);
            if constexpr (ELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) out[e] *= dact[e];
            }
            if (prm.g_x) store_elems<T, 4>(prm.g_x + (row0 + px) * C + 4 * cg, out);
        }
        __syncthreads();                        // the stage is refilled by the prefetch of the next iteration
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    if (!prm.g_w) return;     // uniform across the grid
    // lanes that share a channel group (fixed xor tree over the pixel slots), then warps, then CTAs
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v = (e & 1) ? hi(acc[t][e >> 1]) : lo(acc[t][e >> 1]);
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

// ------------------------------------------------------------------------------------------------
// Forward of the same layer with its neighbours folded in (bts_decoder.py:100-103):
//     iconv1           = Conv2D(F/16, 3, activation='elu')(concat1)      # :100  -- only the ACTIVATION is taken over
//     depth_est_scaled = Conv2D(1, 3, padding='same', use_bias=False, activation='sigmoid')(iconv1)      # :102
//     depth_est        = depth_est_scaled * max_depth                                                     # :103
// The library path is four passes over full-resolution maps: ELU (read + write of the (B,H,W,F/16) map), a layout
// conversion of that map (cuDNN has no NHWC kernel for ONE output channel), the convolution, and the sigmoid -- about
// 1.1 ms at B = 32, 480x640 for a layer whose floor is 0.1 ms (the raw conv output read once, one float per pixel written).
// Here x is read ONCE:
//   phase 1  every pixel of a 32x32 output tile plus its one-pixel halo gets its nine per-tap dot products
//            P[t][q] = sum_c elu(x[q][c]) * w[t][c], stored in shared memory, tap-major.  Pixels outside the image give
//            P = 0 (padding='same' pads the ACTIVATED map with zeros; elu(0) = 0).
//   phase 2  y[p] = sum_t P[t][p + t] in tap order, then the optional sigmoid * max_depth, one coalesced store per row.
// Two phase-1 variants share phase 2:
//   depthconv_fwd_kernel      (this one; tuning key 9 = 1) the first version, kept as the measured comparison: FP32 pipe, four
//                             lanes share a pixel (C/4 channels each: a warp instruction fetches whole 32-byte sectors), the
//                             lane's 9 x C/4 weights stay in registers as packed pairs (FFMA2), the four partial sums are
//                             combined by a fixed xor tree (reduce-scatter: a lane ends with the taps it stores).
//                             Exact float32 FMAs; bound by instruction issue (435 us at B = 32, 480x640, C = 32).
//   depthconv_fwd_mma_kernel  (below; the default) the contraction on the tensor cores with the 3xTF32 split and a cp.async
//                             ring for the inputs (340 us).
// The ELU inside the sum is x > 0 ? x : ex2(x*log2e) - 1 (4 instructions per element instead of the 13 of the
// expm1 form used where ELU values are an OUTPUT): its absolute error (<= 2.4e-7 per activation) enters a 9*C-term sum
// of |w| ~ 0.1 products, i.e. < 1e-6 relative to the logit's scale.
// Algorithmic bytes per pixel: (C + 1) * sizeof(T).
// ------------------------------------------------------------------------------------------------
constexpr int kDfTile = 32;                        // output tile edge
constexpr int kDfHalo = kDfTile + 2;
constexpr int kDfNQ = kDfHalo * kDfHalo;           // 1156 halo pixels; 1156 % 32 == 4, so the four tap rows a warp stores
constexpr int kDfThreads = 256;                    //   (taps b, b+2, b+4, b+6 for 8 consecutive pixels) fall on 32 distinct banks

template <typename T> struct DepthConvFwdParams {
    const T *x;          // (B,H,W,C) contiguous
    const float *w;      // [9][C]
    T *y;                // (B,H,W) contiguous
    uint32_t H, W, tiles_x, items;
    FastDiv div_tx, div_ty;
    int act_out;         // 1: sigmoid(y) * out_scale
    float out_scale;
};

// channel of element j (0 <= j < C/4) of lane-in-pixel s: float32 lanes interleave 16-byte chunks (chunk k of lane s =
// channels [16k + 4s, +4): each load instruction covers 64 contiguous bytes of the pixel), bfloat16 lanes own C/4
// contiguous channels (8 or 16 bytes)
template <typename T, int C> __device__ __forceinline__ int dcf_channel(int s, int j) {
    if constexpr (sizeof(T) == 4) return (j >> 2) * 16 + 4 * s + (j & 3);
    else return (C / 4) * s + j;
}
template <typename T, int C> __device__ __forceinline__ void dcf_load(const T *px, int s, float (&v)[C / 4]) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int k = 0; k < C / 16; ++k) {
            float t[4];
            load_elems<T, 4>(px + 16 * k + 4 * s, t);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[4 * k + e] = t[e];
        }
    } else {
        load_elems<T, C / 4>(px + (C / 4) * s, v);
    }
}

__device__ __forceinline__ float elu_in_sum(float x, float em1) { return x > 0.0f ? x : em1; }

// phase 2: thread -> column (threadIdx & 31), rows (threadIdx >> 5) + 8 k;  y[p] = sum_t P[t][p + t] in tap order
template <typename T> __device__ __forceinline__ void dcf_phase2(const DepthConvFwdParams<T> &prm, const float *P, uint32_t b, int y0, int x0) {
    constexpr int NW = kDfThreads / 32;
    const int col = threadIdx.x & 31;
    const int gx = x0 + 1 + col;
#pragma unroll
    for (int k = 0; k < kDfTile / NW; ++k) {
        const int row = (threadIdx.x >> 5) + NW * k;
        const int gy = y0 + 1 + row;
        if (gx < (int)prm.W && gy < (int)prm.H) {
            float acc = 0.0f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) acc += P[(dy * 3 + dx) * kDfNQ + (row + dy) * kDfHalo + col + dx];
            if (prm.act_out) acc = tail_sigmoid(acc) * prm.out_scale;
            store1(prm.y + ((size_t)b * prm.H + gy) * prm.W + gx, acc);
        }
    }
}

template <typename T, int C, bool ELU>
__global__ void __launch_bounds__(kDfThreads, 2) depthconv_fwd_kernel(const __grid_constant__ DepthConvFwdParams<T> prm) {
    constexpr int CPL = C / 4;                 // channels per lane: 8 or 4
    constexpr int NP2 = CPL / 2;               // packed pairs
    constexpr int NW = kDfThreads / 32;
    constexpr int NPASS = (kDfNQ + 7) / 8;     // warp passes of 8 halo pixels
    __shared__ float P[9 * kDfNQ];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int s = lane & 3, pl = lane >> 2;
    const bool h2 = s & 2, h1 = s & 1;
    const int tap0 = (h2 ? 4 : 0) + (h1 ? 2 : 0);   // after the reduce-scatter the lane holds taps tap0, tap0 + 1

    F2 wr[9][NP2];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < NP2; ++j)
            wr[t][j] = f2(__ldg(prm.w + t * C + dcf_channel<T, C>(s, 2 * j)), __ldg(prm.w + t * C + dcf_channel<T, C>(s, 2 * j + 1)));

    for (uint32_t item = blockIdx.x; item < prm.items; item += gridDim.x) {
        uint32_t rest, tx, b, ty;
        prm.div_tx.divmod(item, rest, tx);
        prm.div_ty.divmod(rest, b, ty);
        const int y0 = (int)(ty * kDfTile) - 1, x0 = (int)(tx * kDfTile) - 1;      // image coordinates of halo pixel (0, 0)
        const T *img = prm.x + (size_t)b * prm.H * prm.W * C;

        // One halo pixel per lane group and pass, software-pipelined: the loads of the warp's next pass are issued before the
        // current one is used (two 32-byte requests per lane in flight; the 72 weight registers leave room for no more).
        auto fetch = [&](int pass, float (&v)[CPL]) -> int {
            const int q = pass * 8 + pl;
            const int hr = q / kDfHalo, hc = q - hr * kDfHalo;
            const int gy = y0 + hr, gx = x0 + hc;
            if (q < kDfNQ && gy >= 0 && gy < (int)prm.H && gx >= 0 && gx < (int)prm.W) {
                dcf_load<T, C>(img + ((size_t)gy * prm.W + gx) * C, s, v);
            } else {
#pragma unroll
                for (int j = 0; j < CPL; ++j) v[j] = 0.0f;      // outside the image: P = 0 (elu(0) = 0); beyond the tile: not stored
            }
            return q;
        };
        float cur[CPL];
        int qc = fetch(wid, cur);
#pragma unroll 2
        for (int pass = wid; pass < NPASS; pass += NW) {
            float nxt[CPL];
            const int qn = fetch(pass + NW, nxt);
            F2 xe[NP2];
#pragma unroll
            for (int j = 0; j < NP2; ++j) {
                float a0 = cur[2 * j], a1 = cur[2 * j + 1];
                if constexpr (ELU) {
                    float t0, t1, e0, e1, m0, m1;
                    unpack(mul2(f2(a0, a1), f2(1.442695040888963407f)), t0, t1);
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
                    unpack(add2(f2(e0, e1), f2(-1.0f)), m0, m1);
                    a0 = elu_in_sum(a0, m0);
                    a1 = elu_in_sum(a1, m1);
                }
                xe[j] = f2(a0, a1);
            }
            float a[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                F2 acc = mul2(xe[0], wr[t][0]);
#pragma unroll
                for (int j = 1; j < NP2; ++j) acc = fma2(xe[j], wr[t][j], acc);
                a[t] = lo(acc) + hi(acc);
            }
            // four lanes -> one: reduce-scatter over xor 2, xor 1 (fixed order), tap 8 by a plain butterfly
            float b4[4], c2[2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float send = h2 ? a[j] : a[j + 4];
                const float keep = h2 ? a[j + 4] : a[j];
                b4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float send = h1 ? b4[j] : b4[j + 2];
                const float keep = h1 ? b4[j + 2] : b4[j];
                c2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
            float a8 = a[8];
            a8 += __shfl_xor_sync(0xffffffffu, a8, 2);
            a8 += __shfl_xor_sync(0xffffffffu, a8, 1);
            if (qc < kDfNQ) {
                P[tap0 * kDfNQ + qc] = c2[0];
                P[(tap0 + 1) * kDfNQ + qc] = c2[1];
                if (s == 0) P[8 * kDfNQ + qc] = a8;
            }
#pragma unroll
            for (int j = 0; j < CPL; ++j) cur[j] = nxt[j];
            qc = qn;
        }
        __syncthreads();
        dcf_phase2<T>(prm, P, b, y0, x0);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Phase 1 on the tensor cores.  ncu on the FP32-pipe kernel above (profiles/r01_depthconv_fwd_f32pipe.md): 24.5 warp
// instructions per pixel, issue slots 52 % busy, FMA 37 % + ALU 36 %, DRAM 46 % -- the per-tap dot products are bound by
// instruction issue (the C -> 9 contraction plus its cross-lane reduction), not by HBM: the case where the contraction
// belongs on the tensor cores.  P[16 pixels x 8 taps] = X[16 x C] * W[C x 8] is one mma.sync.m16n8k8 per 8 channels;
// float32 accuracy is kept by the 3xTF32 split (x = x_hi + x_lo, w = w_hi + w_lo, P = x_lo w_hi + x_hi w_lo + x_hi w_hi;
// the dropped x_lo w_lo term and the truncation of the low parts are <= 2^-20 relative per product; accumulation is
// float32).  bfloat16 inputs are exact TF32 values and need no x_lo term.  The ninth tap stays on the FP32 pipe (a
// second n-tile would spend 12 more MMAs on one column).  The contraction index is order-free, so k-slot (j, tig) /
// (j, tig + 4) of the fragments is mapped to the lane's OWN channels: a quad reads the pixel's 128 bytes with whole-sector
// loads and no shuffle is needed anywhere -- the D fragment already has lane (g, tig) holding taps 2 tig, 2 tig + 1 of
// pixels g and g + 8.  Fixed order, bit-reproducible.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Asynchronous global -> shared copies (LDGSTS): a lane copies ITS OWN 8 / 16 bytes of a pixel into a lane-private slot
// and reads the same slot back later, so no barrier is involved (cp.async.wait_group is per thread) and the LDS is
// conflict-free by construction; src_bytes = 0 zero-fills the slot (padding='same').
template <int BYTES> __device__ __forceinline__ void cp_async(uint32_t dst_smem, const void *src, uint32_t src_bytes) {
    // .cg (L1 bypass): measured 340 us against 352 us for .ca at B = 32, 480x640, C = 32 -- although the bypassing form sends
    // each lane's 16 bytes to L2 as its own sector request (ncu: 2.8x the L2 read sectors), L2 is not the limiter here
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int V> struct IntC1 {
    static constexpr int value = V;
};

#ifndef BTSLPG_DCF_MMA_MINB
#define BTSLPG_DCF_MMA_MINB 2
#endif
#ifndef BTSLPG_DCF_STAGES
#define BTSLPG_DCF_STAGES 4            // passes in the per-thread copy ring (STAGES - 1 in flight while one is consumed)
#endif
template <typename T, int C> struct DcfMmaCfg {
    static constexpr int kLaneBytes = (C / 4) * (int)sizeof(T);          // bytes of a pixel one lane owns: 32, 16 or 8
    static constexpr int kChunks = kLaneBytes > 16 ? kLaneBytes / 16 : 1;
    static constexpr int kCopyBytes = kLaneBytes >= 16 ? 16 : 8;
    static constexpr int kStageBytes = 2 * kChunks * kDfThreads * 16;    // two pixels per lane and pass, 16-byte slots
    static constexpr int kPBytes = 9 * kDfNQ * 4;
    static constexpr int kSmemBytes = kPBytes + BTSLPG_DCF_STAGES * kStageBytes;
    static_assert(kPBytes % 16 == 0, "the copy ring must stay 16-byte aligned");
};

template <typename T, int C, bool ELU>
__global__ void __launch_bounds__(kDfThreads, BTSLPG_DCF_MMA_MINB) depthconv_fwd_mma_kernel(const __grid_constant__ DepthConvFwdParams<T> prm) {
    using Cfg = DcfMmaCfg<T, C>;
    constexpr int CPL = C / 4;                 // channels per lane: a quad covers the pixel
    constexpr int NJ = CPL / 2;                // k-steps of 8 channels (2 per lane)
    constexpr bool SPLIT_X = sizeof(T) == 4;
    constexpr int NW = kDfThreads / 32;
    constexpr int NPASS = (kDfNQ + 15) / 16;   // warp passes of 16 halo pixels
    constexpr int D = BTSLPG_DCF_STAGES;
    extern __shared__ __align__(16) unsigned char dcf_smem[];
    float *const P = reinterpret_cast<float *>(dcf_smem);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane >> 2, tig = lane & 3;
    float *const Pl = P + (2 * tig) * kDfNQ;   // the lane's tap rows 2 tig, 2 tig + 1
    const uint32_t ring = smem_u32(dcf_smem + Cfg::kPBytes) + threadIdx.x * 16;       // this thread's slot 0 of stage 0
    const unsigned char *const ring_g = dcf_smem + Cfg::kPBytes + threadIdx.x * 16;

    // B fragments of taps 0..7 (n = g): k rows tig and tig + 4 of k-step j <-> channels dcf_channel(tig, 2j), (tig, 2j + 1)
    uint32_t bh[NJ][2], bl[NJ][2];
    F2 w8[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float w = __ldg(prm.w + g * C + dcf_channel<T, C>(tig, 2 * j + h));
            bh[j][h] = to_tf32(w);
            bl[j][h] = __float_as_uint(w - __uint_as_float(bh[j][h]));
        }
        w8[j] = f2(__ldg(prm.w + 8 * C + dcf_channel<T, C>(tig, 2 * j)), __ldg(prm.w + 8 * C + dcf_channel<T, C>(tig, 2 * j + 1)));
    }
    const int rowC = (int)prm.W * C;           // elements per image row
    // byte offset of the lane's first chunk inside a pixel: float32 lanes interleave 16-byte chunks (see dcf_channel)
    const int lane_off = sizeof(T) == 4 ? 16 * tig : Cfg::kLaneBytes * tig;

    for (uint32_t item = blockIdx.x; item < prm.items; item += gridDim.x) {
        uint32_t rest, tx, b, ty;
        prm.div_tx.divmod(item, rest, tx);
        prm.div_ty.divmod(rest, b, ty);
        const int y0 = (int)(ty * kDfTile) - 1, x0 = (int)(tx * kDfTile) - 1;      // image coordinates of halo pixel (0, 0)
        // a tile whose halo lies inside the image (most of them) runs the loop without any bounds logic
        const bool interior = y0 >= 0 && x0 >= 0 && y0 + kDfHalo <= (int)prm.H && x0 + kDfHalo <= (int)prm.W;     // CTA-uniform
        const unsigned char *img = reinterpret_cast<const unsigned char *>(prm.x + (size_t)b * prm.H * prm.W * C) + lane_off;

        // copy the lane's share of halo pixel q (row-major in the 34 x 34 halo) into slot `sel` (0: D rows g, 1: rows g + 8)
        // of stage `st`.  The address is clamped into the image (always legal); padding pixels are zero-filled by the
        // copy itself (padding='same' pads the ACTIVATED map; elu(0) = 0).  q >= kDfNQ: past the tile, copied, never stored.
        auto issue = [&](auto interior_tag, int st, int sel, int q) {
            constexpr bool INTERIOR = decltype(interior_tag)::value != 0;
            const int hr = min(q / kDfHalo, kDfHalo - 1), hc = q - (q / kDfHalo) * kDfHalo;
            int gy = y0 + hr, gx = x0 + hc;
            uint32_t nbytes = Cfg::kCopyBytes;
            if constexpr (!INTERIOR) {
                if (gy < 0 || gy >= (int)prm.H || gx < 0 || gx >= (int)prm.W) nbytes = 0;
                gy = max(0, min(gy, (int)prm.H - 1));
                gx = max(0, min(gx, (int)prm.W - 1));
            }
            const unsigned char *src = img + (size_t)(gy * rowC + gx * C) * sizeof(T);
            const uint32_t dst = ring + st * Cfg::kStageBytes + sel * (Cfg::kChunks * kDfThreads * 16);
#pragma unroll
            for (int k = 0; k < Cfg::kChunks; ++k) cp_async<Cfg::kCopyBytes>(dst + k * (kDfThreads * 16), src + 64 * k, nbytes);
        };
        auto take = [&](int st, int sel, float (&v)[CPL]) {
            const unsigned char *src = ring_g + st * Cfg::kStageBytes + sel * (Cfg::kChunks * kDfThreads * 16);
            if constexpr (sizeof(T) == 4) {
#pragma unroll
                for (int k = 0; k < Cfg::kChunks; ++k) {
                    const float4 f = *reinterpret_cast<const float4 *>(src + k * (kDfThreads * 16));
                    v[4 * k] = f.x; v[4 * k + 1] = f.y; v[4 * k + 2] = f.z; v[4 * k + 3] = f.w;
                }
            } else if constexpr (CPL == 8) {
                const uint4 u = *reinterpret_cast<const uint4 *>(src);
                v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
                v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
            } else {
                const uint2 u = *reinterpret_cast<const uint2 *>(src);
                v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
            }
        };
        auto activate = [&](float (&v)[CPL]) {
            if constexpr (ELU) {
#pragma unroll
                for (int j = 0; j < CPL; j += 2) {
                    float t0, t1, e0, e1, m0, m1;
                    unpack(mul2(f2(v[j], v[j + 1]), f2(1.442695040888963407f)), t0, t1);
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
                    unpack(add2(f2(e0, e1), f2(-1.0f)), m0, m1);
                    v[j] = elu_in_sum(v[j], m0);
                    v[j + 1] = elu_in_sum(v[j + 1], m1);
                }
            }
        };
        // one pass: halo pixel qa -> D rows g, qa + 8 -> D rows g + 8
        auto process = [&](float (&ca)[CPL], float (&cb)[CPL], int qa) {
            activate(ca);
            activate(cb);
            // three independent accumulator chains (x_hi w_hi, x_lo w_hi, x_hi w_lo): NJ dependent MMAs each
            float big[4] = {0.0f, 0.0f, 0.0f, 0.0f}, sx[4] = {0.0f, 0.0f, 0.0f, 0.0f}, sw[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            F2 t8a, t8b;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                uint32_t ah[4], al[4];
                const float xs[4] = {ca[2 * j], cb[2 * j], ca[2 * j + 1], cb[2 * j + 1]};   // a0 (g, tig), a1 (g+8, tig), a2 (g, tig+4), a3 (g+8, tig+4)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if constexpr (SPLIT_X) {
                        ah[e] = __float_as_uint(xs[e]) & 0xffffe000u;          // truncation: x_lo < 2^-10 |x|, itself cut to TF32 by the MMA
                        al[e] = __float_as_uint(xs[e] - __uint_as_float(ah[e]));
                    } else {
                        ah[e] = __float_as_uint(xs[e]);
                    }
                }
                if constexpr (SPLIT_X) mma_tf32(sx, al, bh[j][0], bh[j][1]);
                mma_tf32(sw, ah, bl[j][0], bl[j][1]);
                mma_tf32(big, ah, bh[j][0], bh[j][1]);
                const F2 xa = f2(ca[2 * j], ca[2 * j + 1]), xb = f2(cb[2 * j], cb[2 * j + 1]);
                t8a = j == 0 ? mul2(xa, w8[0]) : fma2(xa, w8[j], t8a);
                t8b = j == 0 ? mul2(xb, w8[0]) : fma2(xb, w8[j], t8b);
            }
            float s8a = lo(t8a) + hi(t8a), s8b = lo(t8b) + hi(t8b);
            s8a += __shfl_xor_sync(0xffffffffu, s8a, 1);
            s8b += __shfl_xor_sync(0xffffffffu, s8b, 1);
            s8a += __shfl_xor_sync(0xffffffffu, s8a, 2);
            s8b += __shfl_xor_sync(0xffffffffu, s8b, 2);
            float d[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) d[e] = SPLIT_X ? big[e] + (sx[e] + sw[e]) : big[e] + sw[e];
            const int qb = qa + 8;
            if (qa < kDfNQ) {
                Pl[qa] = d[0];
                Pl[kDfNQ + qa] = d[1];
                if (tig == 0) P[8 * kDfNQ + qa] = s8a;
            }
            if (qb < kDfNQ) {
                Pl[qb] = d[2];
                Pl[kDfNQ + qb] = d[3];
                if (tig == 0) P[8 * kDfNQ + qb] = s8b;
            }
        };

        // per-thread copy ring: D - 1 passes in flight while one is consumed; nothing is held in registers meanwhile
        auto run = [&](auto interior_tag) {
            int qi = wid * 16 + g;                    // next pass to issue (pixel of D row g; row g + 8 is qi + 8)
            int qc = qi;                              // next pass to consume
#pragma unroll
            for (int st = 0; st < D - 1; ++st) {
                issue(interior_tag, st, 0, qi);
                issue(interior_tag, st, 1, qi + 8);
                cp_async_commit();
                qi += NW * 16;
            }
            int st_in = D - 1, st_out = 0;
#pragma unroll 1
            for (int pass = wid; pass < NPASS; pass += NW) {
                issue(interior_tag, st_in, 0, qi);    // unconditional: past the warp's last pass it re-reads a clamped address
                issue(interior_tag, st_in, 1, qi + 8);
                cp_async_commit();
                qi += NW * 16;
                st_in = st_in + 1 == D ? 0 : st_in + 1;
                cp_async_wait<D - 1>();               // the copies of the pass consumed now have landed
                float ca[CPL], cb[CPL];
                take(st_out, 0, ca);
                take(st_out, 1, cb);
                st_out = st_out + 1 == D ? 0 : st_out + 1;
                process(ca, cb, qc);
                qc += NW * 16;
            }
        };
        if (interior) run(IntC1<1>{}); else run(IntC1<0>{});
        cp_async_wait<0>();                       // drain before the ring is reused by the next tile
        __syncthreads();
        dcf_phase2<T>(prm, P, b, y0, x0);
        __syncthreads();
    }
}

}  // namespace btslpg
