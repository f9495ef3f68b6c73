// iconv_kernels.cuh -- iconv1 of the BTS decoder as a tcgen05 implicit GEMM that reads the concat's SOURCES (sm_100a).
//
// Replaces, in inference (SURVEY 8(a) a10 + 8(f) N1; bts_decoder.py:98-100):
//     upconv1 = Conv2D(F/16, 3, activation='elu')(upsample1)                       # :98  only the ACTIVATION
//     concat1 = Concatenate(axis=3)([upconv1, depth_2x2_scaled, depth_4x4_scaled, depth_8x8_scaled])   # :99
//     iconv1  = Conv2D(F/16, 3, padding='same', use_bias=False, activation='elu')(concat1)             # :100 the convolution
// concat1 (35 or 19 channels, 1.38 GB at B = 32, 480x640) is never written: the kernel stages elu(upconv1) and the three
// LPG planes straight from their own buffers.
//
// GEMM view: every staged INPUT row R feeds the three output rows R-1, R, R+1 at once:
//     D[128 pixels][3*NF] += A_R[128][K] * B[K][3*NF],   K = 3 column taps x (NF + 3 -> padded) channels,
// where the N dimension is (kernel row ky, output channel): columns [ky] of the product belong to output row R + 1 - ky.
//   * A: a ring of staged input rows in shared memory, each row stored as planes of 4-channel chunks:
//        row[chunk c][position p][4 floats], position p <-> image column x0 - 1 + p.  This is the tcgen05 K-major
//        no-swizzle canonical layout with SBO = 128 B (8 consecutive positions = one core matrix) and LBO = the plane
//        stride, so the A operand of column tap kx is the SAME data addressed 16 kx bytes further: no im2col copy.
//   * B: the Keras HWIO kernel re-laid in shared memory once per CTA as [kx][chunk][(2 - ky) * NF + n][4]
//        (K-major, LBO = 3*NF*16, SBO = 128); a sub-range of ky at the top / bottom of a segment is a start offset.
//   * D: TMEM holds kIcAcc rotating accumulators of NF columns, one per output row in flight; an MMA with N = 3*NF writes
//        three neighbouring ones.  They are kept ZEROED by the epilogue (tcgen05.st after draining), so every MMA
//        accumulates and one instruction can serve rows in different stages of completion.
// Why this shape: the first formulation (one M128 x N32 tile per output row, nine taps = 45 MMAs that each re-read a 4 KB A
// tile from shared memory) measured 84 cycles per MMA whatever N was (45 MMAs -> 3700 cycles per row, 27 -> 2300; independent
// accumulation chains changed nothing): the tensor core's shared-memory operand reads set the pace, not its arithmetic
// (16 cycles for 128 x 32 x 8) -- profiles/r02_iconv1_tcgen05.md.  Widening N to 3*NF reads each staged row once per column
// tap instead of once per tap: 15 MMAs and ~105 KB of operand reads per row instead of 45 and ~225 KB.
// Warp roles (one CTA of 32 warps per SM): warp 0 issues tcgen05.mma; warps 1-11 copy upconv1 / the planes into the chunk
// planes of ring rows with cp.async (LDGSTS, zero-fill outside the image) as far ahead as the ring allows; warps 12-27 apply
// the ELU and round to TF32 in place once a row has landed (generic proxy -> fence.proxy.async -> mbarrier); warps 28-31
// read TMEM (tcgen05.ld 32x32b), re-zero the accumulator, apply the optional output ELU and store NHWC rows.  mbarrier
// pipelines: row landed (cp.async.mbarrier.arrive), row full, row empty (tcgen05.commit), accumulator full / empty.
//
// Arithmetic: TF32 operands (both rounded to nearest, cvt.rna), float32 accumulation in TMEM -- the precision of the library
// path this replaces (cuDNN under torch's default allow_tf32, TensorFlow's default on Ampere and later).  Stated tolerance
// against the float64 oracle: 3e-3 of the output's largest magnitude (tests/test_iconv_gpu.py).  Fixed order: bit-reproducible.
// Bytes per output pixel: NF*4 (upconv1) + 12 (planes) read, NF*4 written: 268 B at NF = 32.
#pragma once

#include "common.cuh"
#include "tail_kernels.cuh"   // ex2_sfu, kLog2e
#include "tma_pipe.cuh"       // smem_u32, mbarrier helpers, fence_proxy_async

namespace btslpg {

constexpr int kIcThreads = 1024;           // 32 warps (64 registers each): warp 0 MMA, 1-11 copy, 12-27 activation, 28-31 epilogue
constexpr int kIcCopyWarps = 11;           // warps 1..11 issue the cp.async copies (they wait on the memory system, not on issue slots;
                                           // three copy warps could not keep enough LDGSTS in flight: 53 cycles per instruction)
constexpr int kIcProdWarp0 = 12;
constexpr int kIcProdWarps = 16;           // 512 threads: 128 positions x 8 chunks = exactly two 16-byte pieces per thread and row
constexpr int kIcEpiWarp0 = 28;            // warps 28..31 (warp % 4 selects the TMEM lane quarter)
constexpr int kIcTW = 126;                 // output columns of a strip (128 positions of a tile minus the two halo columns)
constexpr int kIcPos = 131;                // staged positions per row: 130 needed; 131 keeps the chunk planes on distinct banks
constexpr int kIcPlane = kIcPos * 16;      // bytes of one chunk plane of a row
constexpr int kIcRing = 7;                 // ring rows: 1-2 under the MMAs, 1 being activated, the rest in flight from global memory
constexpr int kIcAcc = 8;                  // rotating TMEM accumulators (output rows in flight)

template <int NF> struct IconvCfg {
    static constexpr int kCin = NF + 3;                          // [upconv1 (NF), d2, d4, d8]
    static constexpr int kChunks = ((kCin + 3) / 4 + 1) / 2 * 2;  // 4-channel chunks per tap, even (one MMA = 2 chunks): 10 / 6
    static constexpr int kMmaPerTap = kChunks / 2;               // MMAs per column tap
    static constexpr int kPlaneChunk = NF / 4;                   // the chunk that holds [d2, d4, d8, 0]
    static constexpr int kRowBytes = kChunks * kIcPlane;
    static constexpr int kWBytes = 9 * kChunks * NF * 16;
    static constexpr int kTmemCols = kIcAcc * NF;                // 256 / 128: a power of two >= 32
    static constexpr int kBarBytes = 384;
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
    unsigned long long *prof;  // optional (tools only): cycles each role spends waiting, summed over CTAs; NULL in normal use
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
// zero N columns of the warp's 32 TMEM lanes (one register feeds every operand)
template <int N> __device__ __forceinline__ void tmem_zero_row(uint32_t taddr);
template <> __device__ __forceinline__ void tmem_zero_row<32>(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,"
        "%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u) : "memory");
}
template <> __device__ __forceinline__ void tmem_zero_row<16>(uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Round to the nearest TF32 (ties away from zero), as cuDNN / cuBLAS do.  The tensor core ignores the low 13 mantissa bits, so
// adding half a TF32 ulp to the bit pattern IS the rounding (one integer add; cvt.rna.tf32 compiles to four instructions).
// A value within half an ulp of FLT_MAX would carry into infinity -- never the case for activations or kernel weights.
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
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
    uint64_t *bar_full = bars, *bar_empty = bars + kIcRing, *bar_land = bars + 2 * kIcRing, *bar_accf = bars + 3 * kIcRing,
             *bar_acce = bars + 3 * kIcRing + kIcAcc;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * kIcRing + 2 * kIcAcc);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- prologue: kernel -> [kx][chunk][(2 - ky) * NF + n][4] (TF32-rounded), ring zeroed (the padding chunk is never written again)
    for (int idx = threadIdx.x; idx < 9 * CH * NF * 4; idx += kIcThreads) {
        const int e = idx & 3, nn = (idx >> 2) % (3 * NF), kc = idx / (4 * 3 * NF), c = kc % CH, kx = kc / CH;
        const int ky = 2 - nn / NF, n = nn % NF, ch = 4 * c + e;
        wsm[idx] = ch < Cfg::kCin ? to_tf32(__ldg(prm.w + ((size_t)(ky * 3 + kx) * Cfg::kCin + ch) * NF + n)) : 0.0f;
    }
    for (int idx = threadIdx.x; idx < kIcRing * Cfg::kRowBytes / 16; idx += kIcThreads)
        reinterpret_cast<float4 *>(ring)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        for (int k = 0; k < kIcRing; ++k) {
            mbar_init(&bar_full[k], kIcProdWarps);
            mbar_init(&bar_empty[k], 1);
            mbar_init(&bar_land[k], kIcCopyWarps * 32);
        }
        for (int k = 0; k < kIcAcc; ++k) { mbar_init(&bar_accf[k], 1); mbar_init(&bar_acce[k], 4); }
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
    if (warp >= kIcEpiWarp0) {                                        // all accumulators start at zero (every MMA accumulates)
        for (int s = 0; s < kIcAcc; ++s) tmem_zero_row<NF>(tmem_base + ((uint32_t)((warp - kIcEpiWarp0) * 32) << 16) + s * NF);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ================= MMA issuer =================
        // The whole warp walks the loop (so every value is warp-uniform and lives in the uniform datapath that UTCHMMA reads
        // its descriptors from); one elected lane issues.  Descriptors are two 32-bit halves: the high half (SBO, version) is a
        // constant, the low half is (address >> 4) | LBO << 16, so stepping to another tap / chunk pair is one integer add.
        constexpr uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);      // + (N >> 3) << 17
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);                  // SBO = 128 B, descriptor version 1
        const uint32_t a_lo0 = (smem_u32(ring) >> 4) | ((uint32_t)(kIcPlane >> 4) << 16);
        const uint32_t b_lo0 = (smem_u32(wsm) >> 4) | ((uint32_t)(3 * NF * 16 >> 4) << 16);
        uint32_t elected;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(elected));
        // the MMAs of one staged row into `cnt` neighbouring accumulators starting at slot s_lo, kernel rows from n' offset nb
        auto issue = [&](uint32_t row_lo, uint32_t s_lo, uint32_t cnt, uint32_t nb) {
            const uint32_t idesc = idesc0 | (((cnt * NF) >> 3) << 17);
            const uint32_t d_addr = tmem_base + s_lo * NF;
            const uint32_t b_lo = b_lo0 + nb * NF;                               // n' rows are 16 bytes apart: (nb * NF * 16) >> 4
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                for (int jj = 0; jj < Cfg::kMmaPerTap; ++jj) {
                    const uint32_t alo = row_lo + (uint32_t)((2 * jj * kIcPlane + kx * 16) >> 4);
                    const uint32_t blo = b_lo + (uint32_t)((((kx * CH + 2 * jj) * 3 * NF) * 16) >> 4);
                    uint64_t ad, bd;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(ad) : "r"(alo), "r"(desc_hi));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(bd) : "r"(blo), "r"(desc_hi));
                    umma_tf32(d_addr, ad, bd, idesc, 1u);
                }
            }
        };
        uint32_t seq = 0, G = 0;
        long long pw0 = 0, pw1 = 0;
        const long long pt0 = prm.prof ? clock64() : 0;
        for (uint32_t item = blockIdx.x; item < prm.items; item += gridDim.x) {
            const IconvItem it = iconv_item(prm, item);
            for (int jr = 0; jr < it.rows + 2; ++jr, ++seq) {
                // staged row jr (image row r0 - 1 + jr) feeds output rows jr - ky, ky = kylo..kyhi, that exist in this segment
                const int kylo = jr - (it.rows - 1) > 0 ? jr - (it.rows - 1) : 0, kyhi = jr < 2 ? jr : 2;
                const uint32_t slot = seq % kIcRing;
                const long long c0 = prm.prof ? clock64() : 0;
                mbar_wait(&bar_full[slot], (seq / kIcRing) & 1);
                const long long c1 = prm.prof ? clock64() : 0;
                if (kylo == 0) {                                                  // first contribution to output row jr: its accumulator must be drained
                    const uint32_t go = G + jr;
                    if (go >= kIcAcc) mbar_wait(&bar_acce[go % kIcAcc], ((go / kIcAcc) + 1) & 1);
                }
                if (prm.prof) { pw0 += c1 - c0; pw1 += clock64() - c1; }
                tc_fence_after();
                if (elected) {
                    const uint32_t row_lo = a_lo0 + slot * (uint32_t)(Cfg::kRowBytes >> 4);
                    const uint32_t cnt = (uint32_t)(kyhi - kylo + 1), s_lo = (G + jr - kyhi) % kIcAcc, nb = (uint32_t)(2 - kyhi);
                    if (s_lo + cnt <= kIcAcc) {
                        issue(row_lo, s_lo, cnt, nb);
                    } else {                                                      // the accumulator ring wraps inside this row's range
                        const uint32_t n1 = kIcAcc - s_lo;
                        issue(row_lo, s_lo, n1, nb);
                        issue(row_lo, 0, cnt - n1, nb + n1);
                    }
                    tc_commit(&bar_empty[slot]);                                  // the staged row is consumed
                    if (kyhi == 2) tc_commit(&bar_accf[(G + jr - 2) % kIcAcc]);   // output row jr - 2 has all three kernel rows
                }
                __syncwarp();
            }
            G += it.rows;
        }
        if (prm.prof && lane == 0) {
            atomicAdd(prm.prof + 0, (unsigned long long)pw0);
            atomicAdd(prm.prof + 1, (unsigned long long)pw1);
            atomicAdd(prm.prof + 2, (unsigned long long)(clock64() - pt0));
        }
    } else if (warp >= 1 && warp <= kIcCopyWarps) {
        // ================= copy warps: global -> chunk planes of the ring (cp.async / LDGSTS) =================
        // upconv1's raw values go from global memory straight into the chunk planes of a ring row -- no registers, zero-fill
        // outside the image ('same' pads the ACTIVATED map with zeros and elu(0) = 0), 4-byte copies for the three LPG planes.
        // These warps run as far ahead as the ring allows and spend their time blocked on the memory system, which costs no
        // issue slots; with the copies issued by the activation warps (previous version) an SM alternated between a copy
        // phase and an arithmetic phase of ~1000 cycles each and used the memory pipe 38 % of the time.
        constexpr int NC = kIcCopyWarps * 32;                         // 96
        constexpr int CU = NF / 4;                                    // chunks of upconv1 channels per position
        static_assert(NC % CU == 0 && (NC / CU) % 2 == 0, "copy mapping: per-thread chunk index and column parity");
        constexpr int PSTEP = NC / CU;                                // positions between a thread's pieces: 12 / 24
        constexpr int NISS = (128 + PSTEP - 1) / PSTEP;               // pieces per thread and row: 11 / 6
        constexpr int NPL = (3 * 128 + NC - 1) / NC;                  // plane elements per thread and row: 2
        const int ctid = threadIdx.x - 32;
        const int cc = ctid % CU, pp0 = ctid / CU;
        const int Hs = prm.H >> 1, Ws = prm.W >> 1;
        const uint32_t ring_addr = smem_u32(ring);
        const long long sstep = prm.a_subpixel ? (long long)(PSTEP / 2) * 4 * NF : (long long)PSTEP * NF;
        uint32_t seq = 0;
        for (uint32_t item = blockIdx.x; item < prm.items; item += gridDim.x) {
            const IconvItem it = iconv_item(prm, item);
            const int npos = it.sw + 2, xb = it.x0 - 1 + pp0;
            int y = it.r0 - 1;
            // everything that depends only on the item: source pointer of the first row, which columns lie inside the image
            const float *src = prm.a_subpixel
                ? prm.a + (((long long)it.b * Hs + (y >> 1)) * Ws * 4 + (y & 1) * 2) * NF + ((long long)(xb >> 1) * 4 + (xb & 1)) * NF + 4 * cc
                : prm.a + (((long long)it.b * prm.H + y) * prm.W + xb) * NF + 4 * cc;
            uint32_t xin = 0, pin = 0;
#pragma unroll
            for (int k = 0; k < NISS; ++k) {
                const int pp = pp0 + k * PSTEP, x = xb + k * PSTEP;
                if (pp < npos && x >= 0 && x < prm.W) xin |= 1u << k;
            }
            const float *pl[NPL];
#pragma unroll
            for (int k = 0; k < NPL; ++k) {
                const int idx = ctid + k * NC, kp = idx >> 7, pp = idx & 127, x = it.x0 - 1 + pp;
                const float *base = kp == 0 ? prm.p0 : (kp == 1 ? prm.p1 : prm.p2);
                pl[k] = base + ((long long)it.b * prm.H + y) * prm.W + x;
                if (idx < 384 && pp < npos && x >= 0 && x < prm.W) pin |= 1u << k;
            }
            for (int jr = 0; jr < it.rows + 2; ++jr, ++seq, ++y) {
                const uint32_t slot = seq % kIcRing;
                if (seq >= kIcRing) mbar_wait(&bar_empty[slot], ((seq / kIcRing) + 1) & 1);
                const bool yin = y >= 0 && y < prm.H;
                const uint32_t row_addr = ring_addr + slot * Cfg::kRowBytes;
                const float *s = src;
                uint32_t dst = row_addr + cc * kIcPlane + pp0 * 16;
#pragma unroll
                for (int k = 0; k < NISS; ++k) {
                    if (pp0 + k * PSTEP < npos) {
                        const bool in = yin && ((xin >> k) & 1u);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(in ? s : prm.a), "r"(in ? 16 : 0) : "memory");
                    }
                    s += sstep;
                    dst += PSTEP * 16;
                }
#pragma unroll
                for (int k = 0; k < NPL; ++k) {                       // the three LPG planes -> lanes 0..2 of chunk [d2, d4, d8, 0]
                    const int idx = ctid + k * NC, kp = idx >> 7, pp = idx & 127;
                    if (idx < 384 && pp < npos) {
                        const bool in = yin && ((pin >> k) & 1u);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(row_addr + Cfg::kPlaneChunk * kIcPlane + pp * 16 + kp * 4),
                                     "l"(in ? pl[k] : prm.p0), "r"(in ? 4 : 0)
                                     : "memory");
                    }
                    pl[k] += prm.W;
                }
                // the row's "landed" barrier completes when every copy thread's copies of this row have arrived
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar_land[slot])) : "memory");
                // one image row down: the sub-pixel layout alternates between the two row-parity halves of a low-res pixel
                if (prm.a_subpixel) src += (y & 1) ? ((long long)Ws * 4 - 2) * NF : 2 * NF;
                else src += (long long)prm.W * NF;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp >= kIcProdWarp0 && warp < kIcProdWarp0 + kIcProdWarps) {
        // ================= activation warps: ELU + TF32 rounding of a landed row, in place =================
        constexpr int NP = kIcProdWarps * 32;
        constexpr int CU = NF / 4;
        constexpr int NACT = (CU * 128 + NP - 1) / NP;                // 16-byte chunks per thread and row: 2 / 1
        const int ptid = threadIdx.x - kIcProdWarp0 * 32;
        long long pw0 = 0, pw3 = 0;
        const long long pt0 = prm.prof ? clock64() : 0;
        uint32_t seq = 0;
        for (uint32_t item = blockIdx.x; item < prm.items; item += gridDim.x) {
            const IconvItem it = iconv_item(prm, item);
            const int npos = it.sw + 2;
            for (int jr = 0; jr < it.rows + 2; ++jr, ++seq) {
                const uint32_t slot = seq % kIcRing;
                unsigned char *row = ring + slot * Cfg::kRowBytes;
                const long long c0 = prm.prof ? clock64() : 0;
                mbar_wait(&bar_land[slot], (seq / kIcRing) & 1);
                const long long c1 = prm.prof ? clock64() : 0;
                float4 v[NACT];
                float4 *q4[NACT];
                bool on[NACT];
#pragma unroll
                for (int k = 0; k < NACT; ++k) {                      // all loads first: consecutive lanes, consecutive positions of one plane
                    const int idx = ptid + k * NP, c2 = idx >> 7, pp = idx & 127;
                    on[k] = idx < CU * 128 && pp < npos;
                    q4[k] = reinterpret_cast<float4 *>(row + c2 * kIcPlane + pp * 16);
                    if (on[k]) v[k] = *q4[k];
                }
                float4 pl;
                float4 *qp = reinterpret_cast<float4 *>(row + Cfg::kPlaneChunk * kIcPlane + ptid * 16);
                if (ptid < npos) pl = *qp;
#pragma unroll
                for (int k = 0; k < NACT; ++k) {
                    if (on[k]) {
                        v[k].x = elu_tf32(v[k].x); v[k].y = elu_tf32(v[k].y); v[k].z = elu_tf32(v[k].z); v[k].w = elu_tf32(v[k].w);
                        *q4[k] = v[k];
                    }
                }
                if (ptid < npos) {
                    pl.x = to_tf32(pl.x); pl.y = to_tf32(pl.y); pl.z = to_tf32(pl.z);
                    *qp = pl;
                }
                fence_proxy_async();                                   // this thread's stores -> visible to the async proxy (tensor core)
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[slot]);
                if (prm.prof) { pw0 += c1 - c0; pw3 += clock64() - c1; }
            }
        }
        if (prm.prof && ptid == 0) {
            atomicAdd(prm.prof + 5, (unsigned long long)pw0);                       // waiting for the copies to land
            atomicAdd(prm.prof + 6, (unsigned long long)pw3);                       // activation pass
            atomicAdd(prm.prof + 7, (unsigned long long)(clock64() - pt0));
        }
    } else if (warp >= kIcEpiWarp0) {
        // ================= epilogue: TMEM -> registers -> NHWC rows =================
        const int wq = warp - kIcEpiWarp0;                            // == warp % 4: TMEM lanes [32 wq, 32 wq + 32)
        uint32_t g = 0;
        long long pw0 = 0;
        const long long pt0 = prm.prof ? clock64() : 0;
        for (uint32_t item = blockIdx.x; item < prm.items; item += gridDim.x) {
            const IconvItem it = iconv_item(prm, item);
            for (int j = 0; j < it.rows; ++j, ++g) {
                const uint32_t acc = g % kIcAcc;
                const long long c0 = prm.prof ? clock64() : 0;
                mbar_wait(&bar_accf[acc], (g / kIcAcc) & 1);
                if (prm.prof) pw0 += clock64() - c0;
                tc_fence_after();
                uint32_t r[NF];
                const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * NF;
                tmem_ld_row<NF>(taddr, r);
                tmem_ld_wait();
                tmem_zero_row<NF>(taddr);                              // leave the accumulator zeroed for its next output row
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_acce[acc]);
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
        if (prm.prof && warp == kIcEpiWarp0 && lane == 0) {
            atomicAdd(prm.prof + 8, (unsigned long long)pw0);                       // waiting for a finished accumulator
            atomicAdd(prm.prof + 9, (unsigned long long)(clock64() - pt0));
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
