// iconv_kernels.cuh -- iconv1 of the BTS decoder as a tcgen05 implicit GEMM that reads the concat's SOURCES (sm_100a).
//
// Replaces, in inference (SURVEY 8(a) a10 + 8(f) N1; bts_decoder.py:98-100):
//     upconv1 = Conv2D(F/16, 3, activation='elu')(upsample1)                       # :98  only the ACTIVATION
//     concat1 = Concatenate(axis=3)([upconv1, depth_2x2_scaled, depth_4x4_scaled, depth_8x8_scaled])   # :99
//     iconv1  = Conv2D(F/16, 3, padding='same', use_bias=False, activation='elu')(concat1)             # :100 the convolution
// concat1 (35 or 19 channels, 1.38 GB at B = 32, 480x640) is never written: the kernel stages elu(upconv1) and the three
// LPG planes straight from their own buffers.
//
// GEMM view per output row segment: D[128 pixels][NF] += A[128][K] * B[K][NF], K = 9 taps x (NF + 3 -> padded) channels.
//   * A: a ring of staged INPUT rows in shared memory, each row stored as planes of 4-channel chunks:
//        row[chunk c][position p][4 floats], position p <-> image column x0 - 1 + p.  This is the tcgen05 K-major
//        no-swizzle canonical layout with SBO = 128 B (8 consecutive positions = one core matrix) and LBO = the plane
//        stride, so the A operand of tap (ky, kx) is the SAME data addressed at (ring row ky, byte offset 16 kx): no im2col
//        copy, each input row is staged once and used by nine taps of three output rows.
//   * B: the Keras HWIO kernel re-laid in shared memory once per CTA as [tap][chunk][n][4] (K-major, LBO = NF*16, SBO = 128).
//   * D: 128 lanes x NF columns of TMEM, two buffers, so the epilogue of row r overlaps the MMAs of row r + 1.
// Warp roles (one CTA per SM, 12 warps): warp 0 issues tcgen05.mma (one thread); warps 1-7 copy upconv1 / the planes into the
// chunk planes with cp.async four rows ahead, then apply the ELU and round to TF32 in place (generic proxy ->
// fence.proxy.async -> mbarrier); warps 8-11 read TMEM
// (tcgen05.ld 32x32b), apply the optional output ELU and store NHWC rows.  mbarrier pipelines: ring row full / empty
// (empty is signalled by tcgen05.commit), accumulator full / empty.
//
// Arithmetic: TF32 operands (both rounded to nearest, cvt.rna), float32 accumulation in TMEM -- the precision of the library
// path this replaces (cuDNN under torch's default allow_tf32, TensorFlow's default on Ampere and later).  Stated tolerance
// against the float64 oracle: 3e-3 of the output's largest magnitude (tests/test_iconv_gpu.py).
// Bytes per output pixel: NF*4 (upconv1) + 12 (planes) read, NF*4 written: 268 B at NF = 32; the MMA floor is
// 45 x 16 = 720 cycles per 126-pixel row against ~1470 cycles of HBM time, so the kernel is meant to be HBM-bound.
#pragma once

#include "common.cuh"
#include "tail_kernels.cuh"   // ex2_sfu, kLog2e
#include "tma_pipe.cuh"       // smem_u32, mbarrier helpers, fence_proxy_async

namespace btslpg {

constexpr int kIcThreads = 384;
constexpr int kIcProdWarps = 7;            // warps 1..7
constexpr int kIcEpiWarp0 = 8;             // warps 8..11 (warp % 4 selects the TMEM lane quarter)
constexpr int kIcTW = 126;                 // output columns of a strip (128 positions of a tile minus the two halo columns)
constexpr int kIcPos = 131;                // staged positions per row: 130 needed; 131 keeps the chunk planes on distinct banks
constexpr int kIcPlane = kIcPos * 16;      // bytes of one chunk plane of a row
constexpr int kIcRing = 8;                 // ring rows: 3 under the MMAs, 1 being activated, kIcAhead in flight
constexpr int kIcAhead = 4;                // rows whose cp.async copies are in flight ahead of the row being activated

template <int NF> struct IconvCfg {
    static constexpr int kCin = NF + 3;                          // [upconv1 (NF), d2, d4, d8]
    static constexpr int kChunks = ((kCin + 3) / 4 + 1) / 2 * 2;  // 4-channel chunks per tap, even (one MMA = 2 chunks): 10 / 6
    static constexpr int kMmaPerTap = kChunks / 2;
    static constexpr int kPlaneChunk = NF / 4;                   // the chunk that holds [d2, d4, d8, 0]
    static constexpr int kRowBytes = kChunks * kIcPlane;
    static constexpr int kWBytes = 9 * kChunks * NF * 16;
    static constexpr int kParts = 3;                             // partial accumulators per output row (one per kernel row ky)
    static constexpr int kBufCols = kParts * NF;                 // TMEM columns of one accumulator buffer
    static constexpr int kTmemCols = 2 * kBufCols <= 128 ? 128 : 256;   // two buffers; power of two
    static constexpr int kBarBytes = 256;
    static constexpr int kSmemBytes = kIcRing * kRowBytes + kWBytes + kBarBytes;
    static_assert(NF == 16 || NF == 32, "iconv1 has F/16 = 16 or 32 filters");
};

struct IconvParams {
    const float *a;            // upconv1's LINEAR output: (B,H,W,NF), or (B,H/2,W/2,4*NF) when a_subpixel
    const float *p0, *p1, *p2; // depth_2x2 / 4x4 / 8x8_scaled, contiguous (B,H,W)
    const float *w;            // Keras HWIO kernel (3,3,NF+3,NF)
    float *out;                // (B,H,W,NF)
    int B, H, W;
    int a_subpixel, act_out;
    int nstrips, strip_w, rows_per_item, nseg;
    uint32_t items;
};

// ---- tcgen05 / mbarrier wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {          // arrives on `bar` when all MMAs issued so far have completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: 8 rows x 16 B core matrices; lbo = bytes between the two K chunks of an MMA, sbo = bytes between 8-row groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);                                               // descriptor version 1 (sm_100)
}
// D[tmem] (+)= A[smem] * B[smem], TF32 operands, float32 accumulator, M = 128
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
template <int N> __device__ __forceinline__ void tmem_ld_row(uint32_t taddr, uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tmem_ld_row<32>(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
template <> __device__ __forceinline__ void tmem_ld_row<16>(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float to_tf32(float x) {                  // round to nearest (ties away), as cuDNN / cuBLAS do
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// ELU(alpha = 1) inside a TF32 product: exp from the SFU (abs. error 1.2e-7, far below TF32's 2^-11)
__device__ __forceinline__ float elu_tf32(float x) { return to_tf32(x > 0.0f ? x : ex2_sfu(x * kLog2e) - 1.0f); }

struct IconvItem {
    int b, x0, sw, r0, rows;
};
__device__ __forceinline__ IconvItem iconv_item(const IconvParams &prm, uint32_t item) {
    IconvItem it;
    const int s = (int)(item % (uint32_t)prm.nstrips);
    const uint32_t q = item / (uint32_t)prm.nstrips;
    const int seg = (int)(q % (uint32_t)prm.nseg);
    it.b = (int)(q / (uint32_t)prm.nseg);
    it.x0 = s * prm.strip_w;
    it.sw = min(prm.strip_w, prm.W - it.x0);
    it.r0 = seg * prm.rows_per_item;
    it.rows = min(prm.rows_per_item, prm.H - it.r0);
    return it;
}

template <int NF>
__global__ void __launch_bounds__(kIcThreads, 1) iconv1_fwd_kernel(const __grid_constant__ IconvParams prm) {
    using Cfg = IconvCfg<NF>;
    constexpr int CH = Cfg::kChunks;
    extern __shared__ __align__(1024) unsigned char ic_smem[];
    unsigned char *ring = ic_smem;
    float *wsm = reinterpret_cast<float *>(ic_smem + kIcRing * Cfg::kRowBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(ic_smem + kIcRing * Cfg::kRowBytes + Cfg::kWBytes);
    uint64_t *bar_full = bars, *bar_empty = bars + kIcRing, *bar_accf = bars + 2 * kIcRing, *bar_acce = bars + 2 * kIcRing + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kIcRing + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- prologue: kernel -> [tap][chunk][n][4] (TF32-rounded), ring zeroed (the padding chunk is never written again)
    for (int idx = threadIdx.x; idx < 9 * CH * NF * 4; idx += kIcThreads) {
        const int e = idx & 3, n = (idx >> 2) % NF, tc = idx / (4 * NF), c = tc % CH, t = tc / CH, ch = 4 * c + e;
        wsm[idx] = ch < Cfg::kCin ? to_tf32(__ldg(prm.w + ((size_t)t * Cfg::kCin + ch) * NF + n)) : 0.0f;
    }
    for (int idx = threadIdx.x; idx < kIcRing * Cfg::kRowBytes / 16; idx += kIcThreads)
        reinterpret_cast<float4 *>(ring)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        for (int k = 0; k < kIcRing; ++k) { mbar_init(&bar_full[k], kIcProdWarps); mbar_init(&bar_empty[k], 1); }
        for (int k = 0; k < 2; ++k) { mbar_init(&bar_accf[k], 1); mbar_init(&bar_acce[k], 4); }
        mbar_fence_init();
    }
    if (warp == 0) {                                                  // one warp allocates the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();                                              // weights / zeros (generic proxy) -> visible to the tensor core's reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= MMA issuer =================
        // The whole warp walks the loop (so every value is warp-uniform and lives in the uniform datapath that UTCHMMA reads
        // its descriptors from); one elected lane issues.  Descriptors are two 32-bit halves: the high half (SBO, version) is a
        // constant, the low half is (address >> 4) | LBO << 16, so stepping to another tap / chunk pair is ONE integer add of a
        // compile-time constant.  (First version: 64-bit descriptor arithmetic and a modulo per MMA on a single lane, ~80 cycles
        // per MMA -- the issue loop, not the tensor pipe (19 % busy) or HBM (22 %), set the pace: 3700 cycles per row.)
        {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NF >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);                  // SBO = 128 B, descriptor version 1
            const uint32_t a_lo0 = (smem_u32(ring) >> 4) | ((uint32_t)(kIcPlane >> 4) << 16);
            const uint32_t b_lo0 = (smem_u32(wsm) >> 4) | ((uint32_t)(NF * 16 >> 4) << 16);
            uint32_t elected;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(elected));
            uint32_t seq = 0, g = 0;
            for (uint32_t item = blockIdx.x; item < prm.items; item += gridDim.x) {
                const IconvItem it = iconv_item(prm, item);
                mbar_wait(&bar_full[seq % kIcRing], (seq / kIcRing) & 1);
                mbar_wait(&bar_full[(seq + 1) % kIcRing], ((seq + 1) / kIcRing) & 1);
                for (int j = 0; j < it.rows; ++j, ++g) {
                    const uint32_t top = seq + j;
                    mbar_wait(&bar_full[(top + 2) % kIcRing], ((top + 2) / kIcRing) & 1);
                    const uint32_t buf = g & 1;
                    if (g >= 2) mbar_wait(&bar_acce[buf], ((g >> 1) + 1) & 1);       // the epilogue has drained this buffer
                    tc_fence_after();
                    if (elected) {
                        // Three independent accumulation chains (one per kernel row ky, each in its own TMEM columns), issued
                        // round-robin: consecutive tcgen05.mma into ONE accumulator complete ~84 cycles apart whatever N is
                        // (measured: 45 MMAs -> 3700 cycles per row, 27 -> 2300), far above the 16-cycle pipe time of a
                        // 128 x 32 x 8 step; independent chains overlap.  The epilogue adds the three partial tiles.
                        const uint32_t d_addr = tmem_base + buf * Cfg::kBufCols;
                        uint32_t row_lo[3];
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) row_lo[ky] = a_lo0 + ((top + ky) % kIcRing) * (uint32_t)(Cfg::kRowBytes >> 4);
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                            for (int jj = 0; jj < Cfg::kMmaPerTap; ++jj) {
#pragma unroll
                                for (int ky = 0; ky < 3; ++ky) {
                                    const uint32_t alo = row_lo[ky] + (uint32_t)((2 * jj * kIcPlane + kx * 16) >> 4);
                                    const uint32_t blo = b_lo0 + (uint32_t)(((((ky * 3 + kx) * CH + 2 * jj) * NF) * 16) >> 4);
                                    uint64_t ad, bd;
                                    asm("mov.b64 %0, {%1, %2};" : "=l"(ad) : "r"(alo), "r"(desc_hi));
                                    asm("mov.b64 %0, {%1, %2};" : "=l"(bd) : "r"(blo), "r"(desc_hi));
                                    umma_tf32(d_addr + ky * NF, ad, bd, idesc, (kx | jj) ? 1u : 0u);
                                }
                            }
                        }
                        tc_commit(&bar_accf[buf]);                      // accumulator ready for the epilogue
                        tc_commit(&bar_empty[top % kIcRing]);           // the top input row is not needed again
                        if (j == it.rows - 1) {
                            tc_commit(&bar_empty[(top + 1) % kIcRing]);
                            tc_commit(&bar_empty[(top + 2) % kIcRing]);
                        }
                    }
                    __syncwarp();
                }
                seq += it.rows + 2;
            }
        }
    } else if (warp <= kIcProdWarps) {
        // ================= producers: stage input rows =================
        // Two passes per row.  (1) cp.async (LDGSTS) copies of 16 bytes move upconv1's raw values from global memory straight
        // into the chunk planes of a ring row kIcAhead rows ahead of the one being finished -- no registers, zero-fill outside
        // the image ('same' pads the ACTIVATED map with zeros and elu(0) = 0), 4-byte copies for the three LPG planes.  (2) When
        // a row has landed, the producers apply ELU + TF32 rounding to it IN PLACE (conflict-free LDS.128 / STS.128) and hand
        // it to the tensor core.  With the loads held in registers (first two versions) a CTA had one row (~16 KB) in flight
        // and ran at 1.0 TB/s: latency-bound.  Here kIcAhead rows (64 KB) are in flight per SM.
        constexpr int NP = kIcProdWarps * 32;
        constexpr int CU = NF / 4;                                    // chunks of upconv1 channels per position
        const int ptid = threadIdx.x - 32;
        const int Hs = prm.H >> 1, Ws = prm.W >> 1;
        const uint32_t ring_addr = smem_u32(ring);

        struct Cursor {                                               // walks the staged rows of this CTA's items in order
            uint32_t item;
            int jr;
            IconvItem it;
            bool valid;
        };
        auto cur_init = [&](Cursor &c) {
            c.item = blockIdx.x;
            c.jr = 0;
            c.valid = c.item < prm.items;
            if (c.valid) c.it = iconv_item(prm, c.item);
        };
        auto cur_next = [&](Cursor &c) {
            if (++c.jr >= c.it.rows + 2) {
                c.item += gridDim.x;
                c.jr = 0;
                c.valid = c.item < prm.items;
                if (c.valid) c.it = iconv_item(prm, c.item);
            }
        };
        auto row_issue = [&](const Cursor &c, uint32_t row_addr) {
            const IconvItem &it = c.it;
            const int y = it.r0 - 1 + c.jr, npos = it.sw + 2;
            const bool yin = y >= 0 && y < prm.H;
            for (int idx = ptid; idx < npos * CU; idx += NP) {        // consecutive lanes: consecutive 16-byte pieces of a pixel
                const int pp = idx / CU, cc = idx % CU, x = it.x0 - 1 + pp;
                const bool in = yin && x >= 0 && x < prm.W;
                const float *src = prm.a;
                if (in)
                    src = prm.a_subpixel
                        ? prm.a + ((((size_t)it.b * Hs + (y >> 1)) * Ws + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1))) * NF + 4 * cc
                        : prm.a + (((size_t)it.b * prm.H + y) * prm.W + x) * NF + 4 * cc;
                const int sz = in ? 16 : 0;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(row_addr + cc * kIcPlane + pp * 16), "l"(src), "r"(sz) : "memory");
            }
            for (int idx = ptid; idx < npos * 3; idx += NP) {         // the three LPG planes -> lanes 0..2 of chunk [d2, d4, d8, 0]
                const int k = idx / npos, pp = idx - k * npos, x = it.x0 - 1 + pp;
                const bool in = yin && x >= 0 && x < prm.W;
                const float *base = k == 0 ? prm.p0 : (k == 1 ? prm.p1 : prm.p2);
                const float *src = in ? base + ((size_t)it.b * prm.H + y) * prm.W + x : base;
                const int sz = in ? 4 : 0;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(row_addr + Cfg::kPlaneChunk * kIcPlane + pp * 16 + k * 4), "l"(src), "r"(sz)
                             : "memory");
            }
        };
        auto row_activate = [&](const Cursor &c, unsigned char *row) {
            const int npos = c.it.sw + 2;
            for (int idx = ptid; idx < CU * 128; idx += NP) {         // consecutive lanes: consecutive positions of one plane
                const int cc = idx >> 7, pp = idx & 127;
                if (pp < npos) {
                    float4 *q4 = reinterpret_cast<float4 *>(row + cc * kIcPlane + pp * 16);
                    float4 v = *q4;
                    v.x = elu_tf32(v.x); v.y = elu_tf32(v.y); v.z = elu_tf32(v.z); v.w = elu_tf32(v.w);
                    *q4 = v;
                }
            }
            if (ptid < npos) {
                float4 *q4 = reinterpret_cast<float4 *>(row + Cfg::kPlaneChunk * kIcPlane + ptid * 16);
                float4 v = *q4;
                v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z);
                *q4 = v;
            }
        };

        Cursor ci, ca;
        cur_init(ci);
        ca = ci;
        uint32_t seq_i = 0, seq_a = 0;
#pragma unroll 1
        for (int d = 0; d < kIcAhead; ++d) {                          // prime: kIcAhead rows in flight (the ring is empty: no waits)
            if (ci.valid) {
                row_issue(ci, ring_addr + (seq_i % kIcRing) * Cfg::kRowBytes);
                ++seq_i;
                cur_next(ci);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        while (ca.valid) {
            if (ci.valid) {
                const uint32_t slot = seq_i % kIcRing;
                if (seq_i >= kIcRing) mbar_wait(&bar_empty[slot], ((seq_i / kIcRing) + 1) & 1);
                row_issue(ci, ring_addr + slot * Cfg::kRowBytes);
                ++seq_i;
                cur_next(ci);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(kIcAhead) : "memory");        // this thread's copies of row seq_a have landed
            asm volatile("bar.sync 1, %0;" ::"n"(kIcProdWarps * 32) : "memory");       // ... and every other producer's
            const uint32_t slot = seq_a % kIcRing;
            row_activate(ca, ring + slot * Cfg::kRowBytes);
            fence_proxy_async();                                       // this thread's stores -> visible to the async proxy (tensor core)
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[slot]);
            ++seq_a;
            cur_next(ca);
        }
    } else {
        // ================= epilogue: TMEM -> registers -> NHWC rows =================
        const int wq = warp - kIcEpiWarp0;                            // == warp % 4: TMEM lanes [32 wq, 32 wq + 32)
        uint32_t g = 0;
        for (uint32_t item = blockIdx.x; item < prm.items; item += gridDim.x) {
            const IconvItem it = iconv_item(prm, item);
            for (int j = 0; j < it.rows; ++j, ++g) {
                const uint32_t buf = g & 1;
                mbar_wait(&bar_accf[buf], (g >> 1) & 1);
                tc_fence_after();
                uint32_t r[NF], s[NF];
                const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + buf * Cfg::kBufCols;
                tmem_ld_row<NF>(taddr, r);
                tmem_ld_row<NF>(taddr + NF, s);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < NF; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(s[e]));     // (ky0 + ky1) + ky2: fixed order
                tmem_ld_row<NF>(taddr + 2 * NF, s);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < NF; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(s[e]));
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_acce[buf]);            // the MMA warp may overwrite this buffer
                const int m = wq * 32 + lane;
                if (m < it.sw) {
                    float *dst = prm.out + (((size_t)it.b * prm.H + (it.r0 + j)) * prm.W + (it.x0 + m)) * NF;
                    if (prm.act_out) {
#pragma unroll
                        for (int e = 0; e < NF; ++e) {
                            const float v = __uint_as_float(r[e]);
                            r[e] = __float_as_uint(v > 0.0f ? v : expm1f(v));
                        }
                    }
#pragma unroll
                    for (int e = 0; e < NF; e += 8) stg<8>(dst + e, r + e);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
    }
}

}  // namespace btslpg
