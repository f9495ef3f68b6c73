// wgrad_kernels.cuh -- weight gradient of the decoder's full-resolution 3x3 convolutions on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, TMA-fed, accumulators in TMEM; sm_100a).  SURVEY 8(f) N1.
//
// bts_decoder.py:96-101: upconv1 = Conv2D(F/16, 3)(upsample(iconv2)) and iconv1 = Conv2D(F/16, 3)(concat1) run at FULL resolution with
// 16-64 channels (and conv block 2 at half resolution).  Their weight gradient  dW[ky][kx][ci][co] = sum_{b,y,x} in[b, y+ky-1, x+kx-1, ci] * g[b, y, x, co]  is a GEMM with a
// tiny output (9*Cin x Cout <= 576 x 32) and a reduction over every pixel of the batch (7-14 million): the library's kernels for it
// (64x64 output tiles) run at 4-9 times the time it takes to read the two operands once (profiles/r01_tail_convs_cudnn_vs_fused.json:
// 2.0-2.6 ms against floors of 0.4-0.6 ms at B = 32, 480x640).
//
// The GEMM here:  D[M = (input row r, ci)][N = (kx, co)] += A[M][K = 8 pixels] * B[K][N]   per 32-channel block of either operand, with
//   A = activations of FOUR consecutive input rows x 32 channels (M = 128), pixels as K: the rows y-1, y, y+1 that pair with output
//       row y are the kernel rows ky = 0, 1, 2 -- three of the four M blocks of one instruction (the fourth accumulates a row pairing
//       that is not part of the convolution and is dropped);
//   B = the gradient of output row y, N = 96 = the three kernel columns kx x 32 output channels: the three N blocks are the SAME staged
//       row started one pixel (128 bytes) apart (descriptor LBO = 128 B), so the gradient is staged once with one halo pixel per side
//       (TMA zero-fills outside the image, which is exactly padding='same').  Staging three shifted copies instead (one instruction
//       of N = 32 per kernel column) measured 463 us against 277 us at B = 32, 416x544, 32 -> 16 channels.
// Both operands are "MN-major" (the channel index is the contiguous one: NHWC as it lies in memory, no transposition anywhere), which
// for 32-bit operands means the 128-byte-span / 32-byte-atom swizzle: TMA writes the tiles in that pattern (CU_TENSOR_MAP_SWIZZLE_128B_
// ATOM_32B) and the shared-memory descriptors name it (layout type 1).  One TMA producer thread, one MMA issuer thread, accumulators
// stay in TMEM for the whole kernel (a persistent CTA per SM); at the end four warps move them to a per-CTA partial, and a second
// kernel adds the partials in a fixed order: deterministic, no atomics.
// Arithmetic: the tensor core reads float32 bit patterns as TF32 (low 13 mantissa bits ignored -- truncation, where cuDNN rounds to
// nearest; both are within TF32's 2^-10), products are exact, accumulation is float32.
// Algorithmic bytes per pixel: (Cin + Cout) * 4 (each operand read once).
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "iconv_kernels.cuh"   // tcgen05 / mbarrier wrappers
#include "tma_pipe.cuh"

namespace btslpg {

constexpr int kWgTW = 16;                   // output columns of a work item
constexpr int kWgThreads = 192;             // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue (TMEM lane quarters 2, 3, 0, 1)

// CINB / COUTB: 32-channel blocks of the input and of the gradient handled by one pass (<= 64 channels each)
template <int CINB, int COUTB, int T_> struct WgradCfg {
    static constexpr int kT = T_;                                              // output rows of a work item (16 when both operands have one block)
    static constexpr int kXTile = kWgTW * 128;                                 // one row of one 32-channel block: 16 pixels x 128 B
    static constexpr int kXBlock = (kT + 3) * kXTile;                          // rows y-1 .. y+T of a block + the row slot the last instruction touches
    static constexpr int kXBytes = CINB * kXBlock;
    static constexpr int kGW = kWgTW + 2;                                      // gradient pixels per staged row: one halo pixel on either side
    static constexpr int kGRow = kGW * 128;
    static constexpr int kGBlock = (kT * kGRow + 2 * 128 + 1023) / 1024 * 1024;     // + the two pixels the shifted N blocks of the last row run over
    static constexpr int kGBytes = COUTB * kGBlock;
    static constexpr int kStageBytes = kXBytes + kGBytes;
    static constexpr int kTxBytes = CINB * (kT + 2) * kXTile + COUTB * kT * kGRow;
    static constexpr int kStages = (220 * 1024) / kStageBytes;
    static constexpr int kAccCols = CINB * COUTB * 96;                         // per (input block, gradient block): 3 kernel columns x 32 channels
    static constexpr int kTmemCols = kAccCols <= 128 ? 128 : kAccCols <= 256 ? 256 : 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;      // + alignment slack + barriers
    static_assert(kStages >= 2, "the pipeline needs two stages");
    static_assert(kAccCols <= 512, "accumulators exceed TMEM");
};

struct WgradParams {
    float *partial;            // [gridDim.x][9 * Cin][Cout]
    int B, H, W, Cin, Cout;    // Cout: all gradient channels (row length of the partials)
    int co0, co_n;             // this pass: gradient channels [co0, co0 + co_n)
    int ci0, ci_n;             //            input channels [ci0, ci0 + ci_n)
    uint32_t items, bands, ctiles;
    FastDiv div_ct, div_band;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// MN-major operand in the 128-byte-span / 32-byte-atom swizzle: 128-byte rows (32 channels of a pixel), 4-row atoms;
// lbo = bytes between 32-channel blocks, sbo = bytes between 4-pixel groups
__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46) | (1ull << 61);
}

// map_x: boxes of (32 channels, kWgTW pixels, kT + 2 rows); map_g: boxes of (32 channels, kWgTW + 2 pixels, kT rows)
template <int CINB, int COUTB, int T_>
__global__ void __launch_bounds__(kWgThreads, 1) conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g,
                                                                      const __grid_constant__ WgradParams prm) {
    using Cfg = WgradCfg<CINB, COUTB, T_>;
    constexpr int T = Cfg::kT;
    extern __shared__ unsigned char wg_smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(wg_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t *full = bars, *empty = bars + Cfg::kStages, *done = bars + 2 * Cfg::kStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * Cfg::kStages + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem_base = smem_u32(smem);

    if (warp == 0) {
        // ================= TMA producer: CINB + COUTB boxes per work item =================
        if (lane == 0) {
            uint32_t k = 0;
            for (uint32_t item = blockIdx.x; item < prm.items; item += gridDim.x, ++k) {
                const uint32_t s = k % Cfg::kStages, use = k / Cfg::kStages;
                if (use > 0) mbar_wait(empty + s, (use - 1) & 1);
                uint32_t rest, ct, b, band;
                prm.div_ct.divmod(item, rest, ct);
                prm.div_band.divmod(rest, b, band);
                const int x0 = (int)ct * kWgTW, y0 = (int)band * T;
                const uint32_t xs = smem_base + s * Cfg::kStageBytes, gs = xs + Cfg::kXBytes;
                mbar_arrive_expect_tx(full + s, Cfg::kTxBytes);
#pragma unroll
                for (int c = 0; c < CINB; ++c) tma_load_4d(xs + c * Cfg::kXBlock, &map_x, prm.ci0 + c * 32, x0, y0 - 1, (int)b, full + s);
#pragma unroll
                for (int cb = 0; cb < COUTB; ++cb) tma_load_4d(gs + cb * Cfg::kGBlock, &map_g, prm.co0 + cb * 32, x0 - 1, y0, (int)b, full + s);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // instruction descriptor: D float32, A / B TF32, both MN-major (bits 15, 16), M = 128, N = 96
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((96u >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t elected;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(elected));
        uint32_t k = 0;
        for (uint32_t item = blockIdx.x; item < prm.items; item += gridDim.x, ++k) {
            const uint32_t s = k % Cfg::kStages, use = k / Cfg::kStages;
            mbar_wait(full + s, use & 1);
            tc_fence_after();
            const uint32_t xs = smem_base + s * Cfg::kStageBytes, gs = xs + Cfg::kXBytes;
            if (elected) {
                for (int y = 0; y < T; ++y)
#pragma unroll
                    for (int kb = 0; kb < kWgTW / 8; ++kb) {
                        const uint32_t acc = (k | y | kb) != 0;
#pragma unroll
                        for (int cb = 0; cb < COUTB; ++cb) {
                            // N block n' starts n' pixels (n' * 128 bytes) further: B[(n', co)][k] = g[x + n' - 1], kernel column kx = 2 - n'
                            const uint64_t bdesc = umma_desc_mn32(gs + cb * Cfg::kGBlock + y * Cfg::kGRow + kb * 1024, 128, 512);
#pragma unroll
                            for (int c = 0; c < CINB; ++c)
                                umma_tf32(tmem_base + (c * COUTB + cb) * 96, umma_desc_mn32(xs + c * Cfg::kXBlock + y * Cfg::kXTile + kb * 1024, Cfg::kXTile, 512),
                                          bdesc, idesc, acc);
                        }
                    }
                tc_commit(empty + s);
            }
            __syncwarp();
        }
        if (elected) tc_commit(done);
        __syncwarp();
    } else {
        // ================= epilogue: TMEM -> per-CTA partial =================
        // D row m = 32 r + channel: r = input row relative to y - 1 = kernel row ky (r = 3: not part of the convolution)
        mbar_wait(done, 0);
        tc_fence_after();
        const int q = warp & 3;                                   // TMEM lane quarter this warp may read = r
        float *dst0 = prm.partial + (size_t)blockIdx.x * 9 * prm.Cin * prm.Cout;
#pragma unroll
        for (int c = 0; c < CINB; ++c) {
            const int ci = c * 32 + lane;
#pragma unroll
            for (int cb = 0; cb < COUTB; ++cb)
#pragma unroll
                for (int n = 0; n < 3; ++n) {
                    uint32_t v[32];
                    tmem_ld_row<32>(tmem_base + ((uint32_t)(q * 32) << 16) + (c * COUTB + cb) * 96 + n * 32, v);
                    tmem_ld_wait();
                    if (q < 3 && ci < prm.ci_n) {
                        float *d = dst0 + ((size_t)(q * 3 + (2 - n)) * prm.Cin + prm.ci0 + ci) * prm.Cout + prm.co0 + cb * 32;
                        const int nco = min(32, prm.co_n - cb * 32);
                        for (int co = 0; co < nco; ++co) d[co] = __uint_as_float(v[co]);
                    }
                }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
    }
}

// partial rows -> g_w in a fixed order: a warp owns 32 consecutive elements and every 8th row (coalesced 128-byte loads, four in
// flight), the eight row slices are combined in slice order through shared memory
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ partial, float *__restrict__ out, uint32_t n, uint32_t rows) {
    __shared__ float comb[8][32];
    const uint32_t lane = threadIdx.x & 31, sl = threadIdx.x >> 5, i = blockIdx.x * 32 + lane;
    float acc = 0.0f;
    if (i < n) {
        uint32_t r = sl;
        for (; r + 24 < rows; r += 32) {
            const float a0 = __ldcg(partial + (size_t)r * n + i), a1 = __ldcg(partial + (size_t)(r + 8) * n + i);
            const float a2 = __ldcg(partial + (size_t)(r + 16) * n + i), a3 = __ldcg(partial + (size_t)(r + 24) * n + i);
            acc += a0; acc += a1; acc += a2; acc += a3;
        }
        for (; r < rows; r += 8) acc += __ldcg(partial + (size_t)r * n + i);
    }
    comb[sl][lane] = acc;
    __syncthreads();
    if (sl == 0 && i < n) {
#pragma unroll
        for (int q = 1; q < 8; ++q) acc += comb[q][lane];
        out[i] = acc;
    }
}

}  // namespace btslpg
