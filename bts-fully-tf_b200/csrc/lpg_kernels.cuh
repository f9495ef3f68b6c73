// lpg_kernels.cuh -- Local Planar Guidance forward / backward for sm_100a.
//
// What the reference does (custom_layers.py:47-56): decode (phi, theta, dist) -> plane
// (n1,n2,n3,n4), nearest-neighbour expand r x r with two K.repeat_elements, multiply with a
// (1,H,W,3) constant of unit pixel directions, sum, add eps, divide.  ~1.1 GB of HBM traffic for
// 44-69 MB of algorithmic bytes at B=32, 480x640.
//
// What these kernels do.  A "group" is PX horizontally adjacent coarse pixels.  One thread owns a
// group and ROWS rows of its r x r patches: ROWS = r (the whole patch) everywhere except the
// float32 r = 8 backward and the bfloat16 r = 8 kernels, where a patch is split over LPP = r/ROWS
// consecutive WARPS of a CTA (warp s takes rows [s*ROWS, (s+1)*ROWS)) so that no thread has to hold
// 64+ gradients in registers.  A thread reads the group's 3*PX coefficients once, decodes the
// planes, and produces its rows; each row of PX*r outputs is one 16- or 32-byte store, so a warp
// writes 256-1024 contiguous bytes per row and instruction.  Loops over the patch are fully
// unrolled and the row index is a compile-time constant (the warp-uniform split index is turned
// into one by a switch), so the direction table is read as constant-bank operands of the packed
// FMAs themselves: no loads, no registers.  The (1,H,W,3) constant of the reference is never
// materialised.  The strided down-sampled copy (bts_decoder.py:81,88) is written from the same
// registers.
//
// Backward keeps the same ownership: the patch reduction is a fixed-order sum
// inside each thread (packed even/odd column sums over the rows, folded once) followed, for
// LPP > 1, by a fixed-order add of the LPP warps' partials exchanged through shared memory:
// deterministic, no atomics.
//
// Arithmetic:  den = w_pq * (a_p*n1 + b_q*n2 + n3) + eps   with (a_p*w_pq, b_q*w_pq, w_pq) the
// float32 unit direction of custom_layers.py:43 -- the reference's sum(pixel_dir_unit * n) + eps
// up to float32 rounding, in 2 FFMA per pixel; out = n4 * rcp(den) (MUFU.RCP, <= 1 ulp).
#pragma once

#include "common.cuh"
#include "lpg_dir_tables.h"

namespace btslpg {

template <int R> struct DirTable;
template <> struct DirTable<2> {
    static __device__ __forceinline__ float off(int p) { return c_off2[p]; }
    static __device__ __forceinline__ float w(int k) { return c_w2[k]; }
    static __device__ __forceinline__ float v(int k) { return c_v2[k]; }
    static __device__ __forceinline__ F2 off2(int q) { return f2(*reinterpret_cast<const float2 *>(&c_off2[q])); }
    static __device__ __forceinline__ F2 w2(int k) { return f2(*reinterpret_cast<const float2 *>(&c_w2[k])); }
    static __device__ __forceinline__ F2 v2(int k) { return f2(*reinterpret_cast<const float2 *>(&c_v2[k])); }
    static __device__ __forceinline__ F2 u2(int k) { return f2(*reinterpret_cast<const float2 *>(&c_u2[k])); }
};
template <> struct DirTable<4> {
    static __device__ __forceinline__ float off(int p) { return c_off4[p]; }
    static __device__ __forceinline__ float w(int k) { return c_w4[k]; }
    static __device__ __forceinline__ float v(int k) { return c_v4[k]; }
    static __device__ __forceinline__ F2 off2(int q) { return f2(*reinterpret_cast<const float2 *>(&c_off4[q])); }
    static __device__ __forceinline__ F2 w2(int k) { return f2(*reinterpret_cast<const float2 *>(&c_w4[k])); }
    static __device__ __forceinline__ F2 v2(int k) { return f2(*reinterpret_cast<const float2 *>(&c_v4[k])); }
    static __device__ __forceinline__ F2 u2(int k) { return f2(*reinterpret_cast<const float2 *>(&c_u4[k])); }
};
template <> struct DirTable<8> {
    static __device__ __forceinline__ float off(int p) { return c_off8[p]; }
    static __device__ __forceinline__ float w(int k) { return c_w8[k]; }
    static __device__ __forceinline__ float v(int k) { return c_v8[k]; }
    static __device__ __forceinline__ F2 off2(int q) { return f2(*reinterpret_cast<const float2 *>(&c_off8[q])); }
    static __device__ __forceinline__ F2 w2(int k) { return f2(*reinterpret_cast<const float2 *>(&c_w8[k])); }
    static __device__ __forceinline__ F2 v2(int k) { return f2(*reinterpret_cast<const float2 *>(&c_v8[k])); }
    static __device__ __forceinline__ F2 u2(int k) { return f2(*reinterpret_cast<const float2 *>(&c_u8[k])); }
};

struct Angles {
    float sp, cp, st, ct;
};

// ------------------------------------------------------------------------------------------------
// sin and cos of a float32 angle, accurate to < 0.6 ulp(1) (7e-8 abs) for |a| <= 1000:
// quadrant reduction k = rint(a*2/pi) by the 1.5*2^23 trick, r = a - k*pi/2 with a two-term
// Cody-Waite pi/2 (exact products inside the FMAs), degree-7 / degree-8 polynomials on |r| <= pi/4
// (coefficients and the float32 error analysis: tools/fit_sincos.py), then the quadrant swap/sign.
// ~21 instructions for both values; CUDA's sincosf costs ~2x that because of its conversions and
// its large-argument path.  Angles outside the fast range (never produced by a sigmoid head) take
// sincosf().
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void quadrant_fix(int n, float sn, float cs, float &sn_out, float &cs_out) {
    const bool swap = n & 1;
    const float so = swap ? cs : sn;
    const float co = swap ? sn : cs;
    sn_out = __int_as_float(__float_as_int(so) ^ ((n << 30) & 0x80000000));
    cs_out = __int_as_float(__float_as_int(co) ^ (((n + 1) << 30) & 0x80000000));
}

// both angles of a coefficient at once: every floating-point step is one packed instruction
__device__ __forceinline__ void sincos_quadrant2(F2 a, Angles &o) {
    const F2 kf = fma2(a, f2(0x1.45f306p-1f), f2(12582912.0f));     // a*2/pi + 1.5*2^23
    const F2 q = add2(kf, f2(-12582912.0f));
    F2 r = fma2(q, f2(-0x1.921fb6p+0f), a);
    r = fma2(q, f2(0x1.777a5cp-25f), r);
    const F2 s = mul2(r, r);
    F2 ps = fma2(f2(-0x1.9ac9e4p-13f), s, f2(0x1.110c2ap-7f));
    ps = fma2(ps, s, f2(-0x1.555552p-3f));
    const F2 sn = fma2(ps, mul2(r, s), r);
    F2 pc = fma2(f2(0x1.9a6f38p-16f), s, f2(-0x1.6c0e08p-10f));
    pc = fma2(pc, s, f2(0x1.55554cp-5f));
    pc = fma2(pc, s, f2(-0.5f));
    const F2 cs = fma2(pc, s, f2(1.0f));
    quadrant_fix(__float_as_int(lo(kf)), lo(sn), lo(cs), o.sp, o.cp);
    quadrant_fix(__float_as_int(hi(kf)), hi(sn), hi(cs), o.st, o.ct);
}

// out-of-range path (huge, inf or NaN inputs): results come back in registers, so taking it does not
// force the fast path's values through local memory
static __device__ __noinline__ float4 decode_angles_slow(float phi, float theta) {
    float4 r;
    sincosf(phi, &r.x, &r.y);
    sincosf(theta, &r.z, &r.w);
    return r;
}

// custom_layers.py:49 -- phi = x0*2*pi ; theta = x1*pi/3, both rounded to float32 exactly as the
// reference rounds them: (x0*2)*pi == x0*(2*pi) bit for bit (scaling by 2 is exact), and x*pi/3 uses
// a correctly rounded division by 3 in three instructions (q = t/3 approx, one exact residual
// step; verified against IEEE division in tools/fit_sincos.py).
//
// SFU = true (bfloat16 I/O only): the four values come from MUFU.SIN / MUFU.COS (sin.approx, absolute
// error <= 2^-20.9 on the fast range), three orders of magnitude below the 2^-9 resolution of the
// bfloat16 outputs -- the bf16 kernels are instruction-issue-bound (half the bytes, same work), and
// the polynomial decode is a third of a short thread's instructions.  float32 never takes this path.
__device__ __forceinline__ float sin_sfu(float x) {
    float r;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float cos_sfu(float x) {
    float r;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <bool SFU = false> __device__ __forceinline__ void decode_angles(float x0, float x1, Angles &a) {
    if constexpr (SFU) {
        const F2 pt = mul2(f2(x0, x1), f2(2.0f * BTSLPG_PI_F, BTSLPG_PI_F / 3.0f));
        const float phi = lo(pt), theta = hi(pt);
        if (fmaxf(fabsf(phi), fabsf(theta)) <= 1000.0f) {
            a.sp = sin_sfu(phi); a.cp = cos_sfu(phi); a.st = sin_sfu(theta); a.ct = cos_sfu(theta);
        } else {
            const float4 r = decode_angles_slow((x0 * 2.0f) * BTSLPG_PI_F, __fdiv_rn(x1 * BTSLPG_PI_F, 3.0f));
            a.sp = r.x; a.cp = r.y; a.st = r.z; a.ct = r.w;
        }
    } else {
        const F2 pt = mul2(f2(x0, x1), f2(2.0f * BTSLPG_PI_F, BTSLPG_PI_F));   // (phi, x1*pi)
        const float phi = lo(pt), t = hi(pt);
        const float q0 = t * 0x1.555556p-2f;
        const float theta = fmaf(fmaf(-3.0f, q0, t), 0x1.555556p-2f, q0);
        if (fmaxf(fabsf(phi), fabsf(theta)) <= 1000.0f) {
            sincos_quadrant2(f2(phi, theta), a);
        } else {   // IEEE division for theta as well (the 3-instruction form assumes no overflow)
            const float4 r = decode_angles_slow((x0 * 2.0f) * BTSLPG_PI_F, __fdiv_rn(t, 3.0f));
            a.sp = r.x; a.cp = r.y; a.st = r.z; a.ct = r.w;
        }
    }
}
// the decode a kernel with element type T uses
template <typename T> __device__ __forceinline__ void decode_angles_for(float x0, float x1, Angles &a) {
    decode_angles<sizeof(T) == 2>(x0, x1, a);
}

// ------------------------------------------------------------------------------------------------
// Direction table of the ROWS patch rows [SUB*ROWS, SUB*ROWS+ROWS) a thread owns.  Every index is a
// compile-time constant after unrolling, so the entries are constant-bank operands of the (packed)
// FMAs: no loads, no registers.  When a patch is split over LPP = R/ROWS threads, SUB is the index
// of the thread's WARP inside its group of LPP warps (warp-uniform, dispatched by a switch), never a
// per-lane value -- a divergent table index would serialise the constant loads.
// ------------------------------------------------------------------------------------------------
template <int R, int ROWS, int SUB> struct Dirs {
    using Tab = DirTable<R>;
    static __device__ __forceinline__ float a(int k) { return Tab::off(SUB * ROWS + k); }
    static __device__ __forceinline__ F2 w2(int k, int q) { return Tab::w2((SUB * ROWS + k) * R + q); }
    static __device__ __forceinline__ F2 v2(int k, int q) { return Tab::v2((SUB * ROWS + k) * R + q); }
    static __device__ __forceinline__ F2 u2(int k, int q) { return Tab::u2((SUB * ROWS + k) * R + q); }
};

// does patch row (SUB*ROWS + k) carry a down-sampled sample (row % D == 0)?
template <int ROWS, int D, int SUB> __device__ __forceinline__ constexpr bool ds_row(int k) {
    return D != 0 && ((SUB * ROWS + k) % (D ? D : 1)) == 0;
}

// call f(IntC<sub>) with `sub` (0 <= sub < N, warp-uniform) turned into a compile-time constant
template <int N, typename F> __device__ __forceinline__ void dispatch_sub(int sub, F &&f) {
    static_assert(N == 1 || N == 2 || N == 4, "patches are split over 1, 2 or 4 warps");
    if constexpr (N == 1) {
        f(IntC<0>{});
    } else if constexpr (N == 2) {
        if (sub == 0) f(IntC<0>{}); else f(IntC<1>{});
    } else {
        switch (sub) {
            case 0: f(IntC<0>{}); break;
            case 1: f(IntC<1>{}); break;
            case 2: f(IntC<2>{}); break;
            default: f(IntC<3>{}); break;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Expand PX decoded planes into rows [SUB*ROWS, SUB*ROWS+ROWS) of their r x r patches and store
// them; the down-sampled copy x[:, ::d, ::d] (bts_decoder.py:81,88) comes from the same registers.
// orow / drow point at patch row 0 of the group (row stride out_sH / ds_sH).
// ------------------------------------------------------------------------------------------------
template <typename T, int R, int PX, int ROWS, int D, int SUB>
__device__ __forceinline__ void lpg_expand_store(const float (&n1)[PX], const float (&n2)[PX], const float (&n3)[PX],
                                                 const float (&n4)[PX], T *orow, uint32_t out_sH, T *drow, uint32_t ds_sH) {
    using Tab = DirTable<R>;
    using Dir = Dirs<R, ROWS, SUB>;
    constexpr int NDS = D ? R / D : 0;
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
        float o[PX * R];
#pragma unroll
        for (int px = 0; px < PX; ++px) {
            const float A = fmaf(Dir::a(k), n1[px], n3[px]);              // a_p*n1 + n3   (rows pair with n1)
            const F2 A2 = f2(A), n2b = f2(n2[px]), n4b = f2(n4[px]);
#pragma unroll
            for (int q = 0; q < R; q += 2) {                             // two adjacent pixels per packed instruction
                const F2 s = fma2(Tab::off2(q), n2b, A2);                // + b_q*n2      (columns pair with n2)
                const F2 den = fma2(Dir::w2(k, q), s, f2(BTSLPG_EPS_F)); // custom_layers.py:55
                unpack(mul2(n4b, rcp2(den)), o[px * R + q], o[px * R + q + 1]);   // custom_layers.py:56
            }
        }
        constexpr int p0 = SUB * ROWS;
        store_elems<T, PX * R>(orow + (size_t)(p0 + k) * out_sH, o);
        if constexpr (D > 0) {
            if (ds_row<ROWS, D, SUB>(k) && drow) {
                float dsv[PX * NDS];
#pragma unroll
                for (int px = 0; px < PX; ++px)
#pragma unroll
                    for (int qq = 0; qq < NDS; ++qq) dsv[px * NDS + qq] = o[px * R + qq * D];
                store_elems<T, PX * NDS>(drow + (size_t)((p0 + k) / D) * ds_sH, dsv);
            }
        }
    }
}

// Gather G = g_full + scatter(g_ds) for rows [SUB*ROWS, SUB*ROWS+ROWS) of the patches of PX coarse pixels.
template <typename T, int R, int PX, int ROWS, int D, int SUB>
__device__ __forceinline__ void lpg_load_patch(const T *grow, uint32_t gf_sH, const T *drow, uint32_t gd_sH, float (&G)[ROWS][PX * R]) {
    constexpr int NDS = D ? R / D : 0;
    constexpr int p0 = SUB * ROWS;
    if (grow) {
#pragma unroll
        for (int k = 0; k < ROWS; ++k) load_elems<T, PX * R>(grow + (size_t)(p0 + k) * gf_sH, G[k]);
    } else {
#pragma unroll
        for (int k = 0; k < ROWS; ++k)
#pragma unroll
            for (int e = 0; e < PX * R; ++e) G[k][e] = 0.0f;
    }
    if constexpr (D > 0) {
        if (drow) {
#pragma unroll
            for (int k = 0; k < ROWS; ++k) {
                if (ds_row<ROWS, D, SUB>(k)) {
                    float t[PX * NDS];
                    load_elems<T, PX * NDS>(drow + (size_t)((p0 + k) / D) * gd_sH, t);
#pragma unroll
                    for (int px = 0; px < PX; ++px)
#pragma unroll
                        for (int qq = 0; qq < NDS; ++qq) G[k][px * R + qq * D] += t[px * NDS + qq];
                }
            }
        }
    }
}

// Partial sums of one thread over its ROWS rows of the patch of coarse pixel `px`:
//   inv = 1/den ; t = G*inv ; y = t*inv ; acc[3] += t ; acc[2] += y*w ; acc[1] += y*v ; acc[0] += y*u
// (u = a_p*w, v = b_q*w, w: the direction table) so that, after the threads of a group are added,
// g1..g3 = -n4 * acc[0..2] and g4 = acc[3].  The four sums stay PACKED (even / odd columns) over the
// whole patch and are folded once at the end: 8 packed instructions + 2 MUFU.RCP per pair of pixels,
// 2 instructions per row, 4 per patch.  Fixed order: rows, then column pairs, then even + odd.
template <int R, int PX, int ROWS, int SUB>
__device__ __forceinline__ void lpg_patch_partial(const float (&G)[ROWS][PX * R], int px, float n1, float n2, float n3, float (&acc)[4]) {
    using Tab = DirTable<R>;
    using Dir = Dirs<R, ROWS, SUB>;
    const F2 n2b = f2(n2);
    F2 r1, r2, r3, r4;
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
        const F2 A2 = f2(fmaf(Dir::a(k), n1, n3));
#pragma unroll
        for (int q = 0; q < R; q += 2) {
            const F2 w = Dir::w2(k, q);
            const F2 inv = rcp2(fma2(w, fma2(Tab::off2(q), n2b, A2), f2(BTSLPG_EPS_F)));
            const F2 t = mul2(f2(G[k][px * R + q], G[k][px * R + q + 1]), inv);
            const F2 y = mul2(t, inv);
            if (k == 0 && q == 0) {        // resolved at compile time (the loops are unrolled): no zero-initialised sums
                r4 = t;
                r3 = mul2(y, w);
                r2 = mul2(y, Dir::v2(k, q));
                r1 = mul2(y, Dir::u2(k, q));
            } else {
                r4 = add2(r4, t);
                r3 = fma2(y, w, r3);
                r2 = fma2(y, Dir::v2(k, q), r2);
                r1 = fma2(y, Dir::u2(k, q), r1);
            }
        }
    }
    acc[0] = lo(r1) + hi(r1);
    acc[1] = lo(r2) + hi(r2);
    acc[2] = lo(r3) + hi(r3);
    acc[3] = lo(r4) + hi(r4);
}

// acc (summed over the whole patch) -> d loss / d (x0, x1, x2)
__device__ __forceinline__ void lpg_finish_grad(const Angles &a, float n4, const float (&acc)[4], float *gout) {
    // explicit rounding steps (no compiler contraction): bit-identical to lpg_finish_grad2 below
    const float m = -n4;
    const float g1 = __fmul_rn(acc[0], m), g1n = __fmul_rn(acc[0], n4);   // g1, -g1
    const float g2 = __fmul_rn(acc[1], m);
    const float g3n = __fmul_rn(acc[2], n4);                              // -g3
    const float gph = __fmul_rn(a.st, fmaf(g1n, a.sp, __fmul_rn(g2, a.cp)));                      // st*(g2*cp - g1*sp)
    const float gth = fmaf(a.ct, fmaf(g2, a.sp, __fmul_rn(g1, a.cp)), __fmul_rn(g3n, a.st));      // ct*(g1*cp + g2*sp) - g3*st
    gout[0] = __fmul_rn(gph, 2.0f * BTSLPG_PI_F);
    gout[1] = __fmul_rn(gth, BTSLPG_PI_F / 3.0f);
    gout[2] = acc[3];
}
// the same for two coarse pixels at once, every step one packed instruction
__device__ __forceinline__ void lpg_finish_grad2(const Angles &a0, const Angles &a1, float n40, float n41, const float (&acc0)[4],
                                                 const float (&acc1)[4], float *gout0, float *gout1) {
    const F2 n4 = f2(n40, n41), m = f2(-n40, -n41);
    const F2 A0 = f2(acc0[0], acc1[0]);
    const F2 g1 = mul2(A0, m), g1n = mul2(A0, n4);                 // g1, -g1
    const F2 g2 = mul2(f2(acc0[1], acc1[1]), m);
    const F2 g3n = mul2(f2(acc0[2], acc1[2]), n4);                 // -g3
    const F2 sp = f2(a0.sp, a1.sp), cp = f2(a0.cp, a1.cp), st = f2(a0.st, a1.st), ct = f2(a0.ct, a1.ct);
    const F2 gph = mul2(st, fma2(g1n, sp, mul2(g2, cp)));          // st*(g2*cp - g1*sp)
    const F2 gth = fma2(ct, fma2(g2, sp, mul2(g1, cp)), mul2(g3n, st));   // ct*(g1*cp + g2*sp) - g3*st
    unpack(mul2(gph, f2(2.0f * BTSLPG_PI_F)), gout0[0], gout1[0]);
    unpack(mul2(gth, f2(BTSLPG_PI_F / 3.0f)), gout0[1], gout1[1]);
    gout0[2] = acc0[3];
    gout1[2] = acc1[3];
}

// ------------------------------------------------------------------------------------------------
// Vectorised kernels.  groups = B*h*(w/PX).  A warp covers 32 consecutive groups; when a patch is
// split (LPP = R/ROWS > 1) LPP consecutive warps of a CTA cover the SAME 32 groups, warp s taking
// rows [s*ROWS, (s+1)*ROWS).  The CTA size is a multiple of 32*LPP.
// ------------------------------------------------------------------------------------------------
template <typename T> struct LpgFwdParams {
    const T *coef;
    T *out;
    T *ds;                 // nullable
    uint32_t out_sB, out_sH; // element strides of out (column stride 1); < 2^31, checked on the host
    uint32_t ds_sB, ds_sH;
    uint32_t groups;
    FastDiv wg, h;          // groups per coarse row, coarse rows per image
};

template <typename T> struct LpgBwdParams {
    const T *coef;
    const T *g_full;  // nullable
    const T *g_ds;    // nullable
    T *g_coef;
    uint32_t gf_sB, gf_sH;
    uint32_t gd_sB, gd_sH;
    uint32_t groups;
    FastDiv wg, h;
};

template <int R, int ROWS> struct Split {
    static constexpr int LPP = R / ROWS;      // warps per group of 32 coarse-pixel groups
    static_assert(R % ROWS == 0 && (LPP == 1 || LPP == 2 || LPP == 4), "bad row split");
};

// `slot` = index of this thread among the threads of its layer
template <int LPP> __device__ __forceinline__ void slot_to_group(uint32_t slot, uint32_t &group, int &sub) {
    if constexpr (LPP == 1) {
        group = slot;
        sub = 0;
    } else {
        const uint32_t warp = slot >> 5;
        sub = (int)(warp % LPP);
        group = (warp / LPP) * 32 + (slot & 31);
    }
}

// decode + expand + store for one thread's share of group `group`, coefficients already in registers
template <typename T, int R, int PX, int ROWS, int D, int SUB>
__device__ __forceinline__ void lpg_fwd_compute(const LpgFwdParams<T> &prm, uint32_t group, const float (&c)[PX * 3]) {
    constexpr int NDS = D ? R / D : 0;
    uint32_t row, jg, b, i;
    prm.wg.divmod(group, row, jg);
    prm.h.divmod(row, b, i);
    T *orow = prm.out + ((size_t)b * prm.out_sB + (size_t)(i * R) * prm.out_sH + jg * (PX * R));
    T *drow = nullptr;
    if constexpr (D > 0) {
        if (prm.ds) drow = prm.ds + ((size_t)b * prm.ds_sB + (size_t)(i * NDS) * prm.ds_sH + jg * (PX * NDS));
    }
    float n1[PX], n2[PX], n3[PX], n4[PX];
#pragma unroll
    for (int px = 0; px < PX; ++px) {
        Angles a;
        decode_angles_for<T>(c[3 * px], c[3 * px + 1], a);
        n1[px] = a.st * a.cp;   // custom_layers.py:50
        n2[px] = a.st * a.sp;
        n3[px] = a.ct;
        n4[px] = c[3 * px + 2];
    }
    lpg_expand_store<T, R, PX, ROWS, D, SUB>(n1, n2, n3, n4, orow, prm.out_sH, drow, prm.ds_sH);
}

// One wave ahead: the forward is store-dominated, and what a CTA waits for at its start is the DRAM latency of
// its few coefficient bytes (ncu: long_scoreboard is the top stall).  Every fourth lane asks L2 for the line
// that the thread `PF` slots further on will read -- roughly the CTAs that become resident when the current
// ones retire -- so that their first loads are L2 hits.  No registers, no dependency, a handful of instructions.
#ifndef BTSLPG_FWD_PF_CTAS
#define BTSLPG_FWD_PF_CTAS 2368          // 16 resident CTAs x 148 SMs; 0 disables the prefetch
#endif
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <typename T, int R, int PX, int ROWS, int D>
__device__ __forceinline__ void lpg_fwd_thread(const LpgFwdParams<T> &prm, uint32_t slot) {
    constexpr int LPP = Split<R, ROWS>::LPP;
    uint32_t group;
    int sub;
    if constexpr (BTSLPG_FWD_PF_CTAS > 0) {
        if ((threadIdx.x & 3) == 0) {
            uint32_t g2;
            int s2;
            slot_to_group<LPP>(slot + BTSLPG_FWD_PF_CTAS * 128u, g2, s2);
            if (g2 < prm.groups && s2 == 0) prefetch_l2(prm.coef + (size_t)g2 * (PX * 3));
        }
    }
    slot_to_group<LPP>(slot, group, sub);
    if (group >= prm.groups) return;
    float c[PX * 3];
    load_elems<T, PX * 3, 4, true>(prm.coef + (size_t)group * (PX * 3), c);
    dispatch_sub<LPP>(sub, [&](auto S) { lpg_fwd_compute<T, R, PX, ROWS, D, decltype(S)::value>(prm, group, c); });
}

// Backward.  With LPP > 1 the LPP warps of a group exchange their partial sums through shared memory
// and warp 0 of the group adds them in a fixed order ((s0+s1)+(s2+s3)): deterministic, no atomics.
// Every thread of the CTA must call this function when LPP > 1 (it contains a CTA barrier).
template <typename T, int R, int PX, int ROWS, int D>
__device__ __forceinline__ void lpg_bwd_thread(const LpgBwdParams<T> &prm, uint32_t slot) {
    constexpr int LPP = Split<R, ROWS>::LPP;
    constexpr int NDS = D ? R / D : 0;
    uint32_t group;
    int sub;
    slot_to_group<LPP>(slot, group, sub);
    const bool active = group < prm.groups;
    if (LPP == 1 && !active) return;

    float G[ROWS][PX * R];
    float c[PX * 3];
    if (active) {
        // issue every load of the thread first (memory-level parallelism), then compute
        uint32_t row, jg, b, i;
        prm.wg.divmod(group, row, jg);
        prm.h.divmod(row, b, i);
        const T *grow = prm.g_full ? prm.g_full + ((size_t)b * prm.gf_sB + (size_t)(i * R) * prm.gf_sH + jg * (PX * R)) : nullptr;
        const T *drow = nullptr;
        if constexpr (D > 0) {
            if (prm.g_ds) drow = prm.g_ds + ((size_t)b * prm.gd_sB + (size_t)(i * NDS) * prm.gd_sH + jg * (PX * NDS));
        }
        load_elems<T, PX * 3, 4, true>(prm.coef + (size_t)group * (PX * 3), c);
        dispatch_sub<LPP>(sub, [&](auto S) { lpg_load_patch<T, R, PX, ROWS, D, decltype(S)::value>(grow, prm.gf_sH, drow, prm.gd_sH, G); });
    } else {
#pragma unroll
        for (int e = 0; e < PX * 3; ++e) c[e] = 0.0f;
#pragma unroll
        for (int k = 0; k < ROWS; ++k)
#pragma unroll
            for (int e = 0; e < PX * R; ++e) G[k][e] = 0.0f;
    }

    Angles a[PX];
    float acc[PX][4];
#pragma unroll
    for (int px = 0; px < PX; ++px) {
        decode_angles_for<T>(c[3 * px], c[3 * px + 1], a[px]);
        dispatch_sub<LPP>(sub, [&](auto S) {
            lpg_patch_partial<R, PX, ROWS, decltype(S)::value>(G, px, a[px].st * a[px].cp, a[px].st * a[px].sp, a[px].ct, acc[px]);
        });
    }
    if constexpr (LPP > 1) {
        __shared__ float4 part[8][PX][32];                 // [warp of the CTA][px][lane]; CTA size <= 256
        const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int px = 0; px < PX; ++px) part[wid][px][lane] = make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
        __syncthreads();
        if (sub != 0 || !active) return;
#pragma unroll
        for (int px = 0; px < PX; ++px) {
            float4 t[LPP];
            t[0] = make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
#pragma unroll
            for (int s2 = 1; s2 < LPP; ++s2) t[s2] = part[wid + s2][px][lane];
            if constexpr (LPP == 2) {
                acc[px][0] = t[0].x + t[1].x; acc[px][1] = t[0].y + t[1].y; acc[px][2] = t[0].z + t[1].z; acc[px][3] = t[0].w + t[1].w;
            } else {
                acc[px][0] = (t[0].x + t[1].x) + (t[2].x + t[3].x);
                acc[px][1] = (t[0].y + t[1].y) + (t[2].y + t[3].y);
                acc[px][2] = (t[0].z + t[1].z) + (t[2].z + t[3].z);
                acc[px][3] = (t[0].w + t[1].w) + (t[2].w + t[3].w);
            }
        }
    }
    float gout[PX * 3];
    if constexpr (PX % 2 == 0) {
#pragma unroll
        for (int px = 0; px < PX; px += 2)
            lpg_finish_grad2(a[px], a[px + 1], c[3 * px + 2], c[3 * px + 5], acc[px], acc[px + 1], &gout[3 * px], &gout[3 * px + 3]);
    } else {
#pragma unroll
        for (int px = 0; px < PX; ++px) lpg_finish_grad(a[px], c[3 * px + 2], acc[px], &gout[3 * px]);
    }
    store_elems<T, PX * 3, 4>(prm.g_coef + (size_t)group * (PX * 3), gout);
}

template <typename T, int R, int PX, int ROWS, int D>
__global__ void __launch_bounds__(256) lpg_fwd_vec_kernel(const __grid_constant__ LpgFwdParams<T> prm) {
    lpg_fwd_thread<T, R, PX, ROWS, D>(prm, blockIdx.x * blockDim.x + threadIdx.x);
}
template <typename T, int R, int PX, int ROWS, int D>
__global__ void __launch_bounds__(256) lpg_bwd_vec_kernel(const __grid_constant__ LpgBwdParams<T> prm) {
    lpg_bwd_thread<T, R, PX, ROWS, D>(prm, blockIdx.x * blockDim.x + threadIdx.x);
}

// Default (PX, ROWS) of the vectorised variants, chosen for the smallest register footprint (occupancy
// is what hides the load latency here; measured in profiles/r01_sweep_*.json):
//   forward   r8: one thread per coarse pixel, all 8 rows (constant-bank weights, 32 registers)
//   backward  r8: 4 lanes x 2 rows + shuffle tree (holding 8 rows of gradients would need 64 registers)
//   r4: one coarse pixel per thread;  r2: two (float32) / four (bfloat16) coarse pixels per thread
template <typename T> __host__ __device__ constexpr int px_max(int r) {
    return sizeof(T) == 4 ? (r == 2 ? 2 : 1) : (r == 2 ? 4 : 2);
}
template <typename T> __host__ __device__ constexpr int rows_default(int r, bool fwd) {
    return r == 8 ? ((fwd && sizeof(T) == 4) ? 8 : 2) : r;
}
template <typename T, int R, bool FWD> struct VecCfg {
    static constexpr int PX = px_max<T>(R);
    static constexpr int ROWS = rows_default<T>(R, FWD);
    static constexpr int LPP = R / ROWS;
};
// threads needed for `groups` groups
__host__ __device__ inline uint32_t threads_for(uint32_t groups, int lpp) {
    return ((groups + 31) / 32) * 32 * lpp;
}

// ------------------------------------------------------------------------------------------------
// The three scales of one decoder in one launch.  Slot s holds the layer with up-ratio 8 >> s
// (slot 0: r = 8 with ds stride 4, slot 1: r = 4 with ds stride 2, slot 2: r = 2, no ds); an empty
// slot owns no blocks.  Because the slot index is static inside each branch, every parameter is a
// constant-bank operand of the instruction that uses it (no indexed LDC, no register copies), and the
// branch itself is block-uniform.  Blocks of the long r = 8 threads come first, the short r = 2
// threads fill the tail of the grid.
// ------------------------------------------------------------------------------------------------
constexpr int kMultiSlots = 3;
__host__ __device__ constexpr int multi_slot(int r) { return r == 8 ? 0 : r == 4 ? 1 : r == 2 ? 2 : -1; }

template <typename T> struct LpgFwdMulti {
    LpgFwdParams<T> layer[kMultiSlots];
    uint32_t block_end[kMultiSlots];  // exclusive prefix of blocks per slot
};
template <typename T> struct LpgBwdMulti {
    LpgBwdParams<T> layer[kMultiSlots];
    uint32_t block_end[kMultiSlots];
};

// MINB = minimum resident CTAs per SM asked of the compiler (caps registers: occupancy is what hides the
// one-shot load latency of these kernels); block size is 128.
constexpr int kMultiThreads = 128;
// float32: 16 / 12 CTAs per SM (<= 32 / 40 registers; measured: profiles/experiments/README.md); bfloat16 threads hold twice the pixels: 12 / 8
#ifndef BTSLPG_FWD_MINB
#define BTSLPG_FWD_MINB 16
#endif
#ifndef BTSLPG_BWD_MINB
#define BTSLPG_BWD_MINB 12
#endif
#ifndef BTSLPG_FWD_MINB_BF16
#define BTSLPG_FWD_MINB_BF16 12
#endif
#ifndef BTSLPG_BWD_MINB_BF16
#define BTSLPG_BWD_MINB_BF16 8
#endif
template <typename T> constexpr int fwd_min_blocks() { return sizeof(T) == 4 ? BTSLPG_FWD_MINB : BTSLPG_FWD_MINB_BF16; }
template <typename T> constexpr int bwd_min_blocks() { return sizeof(T) == 4 ? BTSLPG_BWD_MINB : BTSLPG_BWD_MINB_BF16; }

template <typename T>
__global__ void __launch_bounds__(kMultiThreads, fwd_min_blocks<T>()) lpg_fwd_multi_kernel(const __grid_constant__ LpgFwdMulti<T> m) {
    const uint32_t blk = blockIdx.x;
    if (blk < m.block_end[0]) {
        lpg_fwd_thread<T, 8, VecCfg<T, 8, true>::PX, VecCfg<T, 8, true>::ROWS, 4>(m.layer[0], blk * kMultiThreads + threadIdx.x);
    } else if (blk < m.block_end[1]) {
        lpg_fwd_thread<T, 4, VecCfg<T, 4, true>::PX, VecCfg<T, 4, true>::ROWS, 2>(m.layer[1], (blk - m.block_end[0]) * kMultiThreads + threadIdx.x);
    } else {
        lpg_fwd_thread<T, 2, VecCfg<T, 2, true>::PX, VecCfg<T, 2, true>::ROWS, 0>(m.layer[2], (blk - m.block_end[1]) * kMultiThreads + threadIdx.x);
    }
}

template <typename T>
__global__ void __launch_bounds__(kMultiThreads, bwd_min_blocks<T>()) lpg_bwd_multi_kernel(const __grid_constant__ LpgBwdMulti<T> m) {
    const uint32_t blk = blockIdx.x;
    if (blk < m.block_end[0]) {
        lpg_bwd_thread<T, 8, VecCfg<T, 8, false>::PX, VecCfg<T, 8, false>::ROWS, 4>(m.layer[0], blk * kMultiThreads + threadIdx.x);
    } else if (blk < m.block_end[1]) {
        lpg_bwd_thread<T, 4, VecCfg<T, 4, false>::PX, VecCfg<T, 4, false>::ROWS, 2>(m.layer[1], (blk - m.block_end[0]) * kMultiThreads + threadIdx.x);
    } else {
        lpg_bwd_thread<T, 2, VecCfg<T, 2, false>::PX, VecCfg<T, 2, false>::ROWS, 0>(m.layer[2], (blk - m.block_end[1]) * kMultiThreads + threadIdx.x);
    }
}

// ------------------------------------------------------------------------------------------------
// Generic kernels: any up-ratio r >= 1, any ds stride d | r, any element strides (e.g. an NHWC
// concat slot whose column stride is the channel count).  One thread per coarse pixel, scalar
// accesses.  Same per-pixel arithmetic as the vectorised kernels (the direction is recomputed
// with the float32 op sequence of custom_layers.py:38-43), so the forward is bit-identical between
// the two paths; the backward differs only in the association of the patch sum.
// ------------------------------------------------------------------------------------------------
template <typename T> struct LpgGenericParams {
    const T *coef;
    int64_t c_sB, c_sH, c_sW, c_sC;
    // forward: out / ds ; backward: o_s* describe g_full and d_s* describe g_ds
    T *out;
    T *ds;
    int64_t o_sB, o_sH, o_sW;
    int64_t d_sB, d_sH, d_sW;
    const T *g_full;
    const T *g_ds;
    T *g_coef;
    int64_t gc_sB, gc_sH, gc_sW, gc_sC;
    int64_t B, h, w;
    int32_t r, d;
};

__device__ __forceinline__ float dir_offset(int p, int r) {
    return ((float)p - (float)(r - 1) * 0.5f) / (float)r;   // (k % r - (r-1)/2) / r, exact in float32
}
__device__ __forceinline__ float dir_w(float a, float b) {
    const float ss = (a * a + b * b) + 1.0f;
    return __fdiv_rn(1.0f, __fsqrt_rn(fmaxf(ss, 1e-12f)));  // l2_normalize: rsqrt(max(sum sq, 1e-12))
}

template <typename T> __global__ void __launch_bounds__(128) lpg_fwd_generic_kernel(const __grid_constant__ LpgGenericParams<T> prm) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= prm.B * prm.h * prm.w) return;
    const int64_t j = idx % prm.w, i = (idx / prm.w) % prm.h, b = idx / (prm.w * prm.h);
    const T *c = prm.coef + b * prm.c_sB + i * prm.c_sH + j * prm.c_sW;
    Angles a;
    decode_angles(load1(c), load1(c + prm.c_sC), a);
    const float n1 = a.st * a.cp, n2 = a.st * a.sp, n3 = a.ct, n4 = load1(c + 2 * prm.c_sC);
    const int r = prm.r, d = prm.d;
    for (int p = 0; p < r; ++p) {
        const float ap = dir_offset(p, r);
        const float A = fmaf(ap, n1, n3);
        for (int q = 0; q < r; ++q) {
            const float bq = dir_offset(q, r);
            const float s = fmaf(bq, n2, A);
            const float den = fmaf(dir_w(ap, bq), s, BTSLPG_EPS_F);
            const float o = n4 * rcp_approx(den);
            store1(prm.out + b * prm.o_sB + (i * r + p) * prm.o_sH + (j * r + q) * prm.o_sW, o);
            if (prm.ds && d > 0 && p % d == 0 && q % d == 0)
                store1(prm.ds + b * prm.d_sB + ((i * r + p) / d) * prm.d_sH + ((j * r + q) / d) * prm.d_sW, o);
        }
    }
}

template <typename T> __global__ void __launch_bounds__(128) lpg_bwd_generic_kernel(const __grid_constant__ LpgGenericParams<T> prm) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= prm.B * prm.h * prm.w) return;
    const int64_t j = idx % prm.w, i = (idx / prm.w) % prm.h, b = idx / (prm.w * prm.h);
    const T *c = prm.coef + b * prm.c_sB + i * prm.c_sH + j * prm.c_sW;
    Angles a;
    decode_angles(load1(c), load1(c + prm.c_sC), a);
    const float n1 = a.st * a.cp, n2 = a.st * a.sp, n3 = a.ct, n4 = load1(c + 2 * prm.c_sC);
    const int r = prm.r, d = prm.d;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p = 0; p < r; ++p) {
        const float ap = dir_offset(p, r);
        const float A = fmaf(ap, n1, n3);
        float r2 = 0.f, r3 = 0.f, r4 = 0.f;
        for (int q = 0; q < r; ++q) {
            const float bq = dir_offset(q, r);
            const float w = dir_w(ap, bq);
            float G = 0.f;
            if (prm.g_full) G = load1(prm.g_full + b * prm.o_sB + (i * r + p) * prm.o_sH + (j * r + q) * prm.o_sW);
            if (prm.g_ds && d > 0 && p % d == 0 && q % d == 0)
                G += load1(prm.g_ds + b * prm.d_sB + ((i * r + p) / d) * prm.d_sH + ((j * r + q) / d) * prm.d_sW);
            const float s = fmaf(bq, n2, A);
            const float inv = rcp_approx(fmaf(w, s, BTSLPG_EPS_F));
            const float u = G * inv;
            const float y = u * inv;
            r4 += u;
            r3 = fmaf(y, w, r3);
            r2 = fmaf(y, bq * w, r2);
        }
        acc[0] = fmaf(ap, r3, acc[0]);
        acc[1] += r2;
        acc[2] += r3;
        acc[3] += r4;
    }
    float gout[3];
    lpg_finish_grad(a, n4, acc, gout);
    T *o = prm.g_coef + b * prm.gc_sB + i * prm.gc_sH + j * prm.gc_sW;
    store1(o, gout[0]);
    store1(o + prm.gc_sC, gout[1]);
    store1(o + 2 * prm.gc_sC, gout[2]);
}

}  // namespace btslpg
