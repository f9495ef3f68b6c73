// lpg_kernels.cuh -- Local Planar Guidance forward / backward for sm_100a.
//
// What the reference does (custom_layers.py:47-56): decode (phi, theta, dist) -> plane
// (n1,n2,n3,n4), nearest-neighbour expand r x r with two K.repeat_elements, multiply with a
// (1,H,W,3) constant of unit pixel directions, sum, add eps, divide.  ~1.1 GB of HBM traffic for
// 44-69 MB of algorithmic bytes at B=32, 480x640.
//
// What these kernels do: ONE thread owns PX horizontally adjacent coarse pixels.  It reads their
// 3*PX coefficients once (vector load), decodes the plane once, and produces the whole r x r
// patch row by row; each row of PX*r outputs is one 16- or 32-byte store, so a warp writes
// 512-1024 contiguous bytes per instruction.  Because loops over the patch are fully unrolled and
// every lane is at the same patch position, the direction table is read as constant-bank operands
// of the FFMAs themselves (no table loads, no materialised (1,H,W,3) constant).  The strided
// down-sampled copy (bts_decoder.py:81,88) is written from the same registers.
//
// Backward keeps the same ownership, so the r x r patch reduction of the gradient (SURVEY 8(a)
// a6) happens entirely inside one thread in a fixed order: deterministic, no atomics, and no
// cross-lane traffic at all.
//
// Denominator:  den = w_pq * (a_p*n1 + b_q*n2 + n3) + eps   with (a_p*w_pq, b_q*w_pq, w_pq) the
// float32 unit direction of custom_layers.py:43 -- the same value as the reference's
// sum(pixel_dir_unit * n) + eps up to float32 rounding (2 FFMA per pixel instead of 3 FMUL/FADD+1).
#pragma once

#include "common.cuh"
#include "lpg_dir_tables.h"

namespace btslpg {

template <int R> struct DirTable;
template <> struct DirTable<2> {
    static __device__ __forceinline__ float off(int p) { return c_off2[p]; }
    static __device__ __forceinline__ float u(int k) { return c_u2[k]; }
    static __device__ __forceinline__ float v(int k) { return c_v2[k]; }
    static __device__ __forceinline__ float w(int k) { return c_w2[k]; }
};
template <> struct DirTable<4> {
    static __device__ __forceinline__ float off(int p) { return c_off4[p]; }
    static __device__ __forceinline__ float u(int k) { return c_u4[k]; }
    static __device__ __forceinline__ float v(int k) { return c_v4[k]; }
    static __device__ __forceinline__ float w(int k) { return c_w4[k]; }
};
template <> struct DirTable<8> {
    static __device__ __forceinline__ float off(int p) { return c_off8[p]; }
    static __device__ __forceinline__ float u(int k) { return c_u8[k]; }
    static __device__ __forceinline__ float v(int k) { return c_v8[k]; }
    static __device__ __forceinline__ float w(int k) { return c_w8[k]; }
};

struct Angles {
    float sp, cp, st, ct;
};

// custom_layers.py:49 -- phi = x0*2*pi ; theta = x1*pi/3 in float32, then full-precision sin/cos
// (no fast-math intrinsics: __sinf is off by 1e-5 near 2*pi).
__device__ __forceinline__ void decode_angles(float x0, float x1, Angles &a) {
    float phi = (x0 * 2.0f) * BTSLPG_PI_F;
    float theta = __fdiv_rn(x1 * BTSLPG_PI_F, 3.0f);
    sincosf(phi, &a.sp, &a.cp);
    sincosf(theta, &a.st, &a.ct);
}

// Expand PX decoded planes into their r x r patches and store them row by row; the strided
// down-sampled copy (bts_decoder.py:81,88  x[:, ::d, ::d]) is written from the same registers.
template <typename T, int R, int PX, int D>
__device__ __forceinline__ void lpg_expand_store(const float (&n1)[PX], const float (&n2)[PX], const float (&n3)[PX],
                                                 const float (&n4)[PX], T *orow, int64_t out_sH, T *drow, int64_t ds_sH) {
    using Tab = DirTable<R>;
    constexpr int NDS = D ? R / D : 0;
#pragma unroll
    for (int p = 0; p < R; ++p) {
        float o[PX * R];
#pragma unroll
        for (int px = 0; px < PX; ++px) {
            const float A = fmaf(Tab::off(p), n1[px], n3[px]);            // a_p*n1 + n3   (rows pair with n1)
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const float s = fmaf(Tab::off(q), n2[px], A);            // + b_q*n2      (columns pair with n2)
                const float den = fmaf(Tab::w(p * R + q), s, BTSLPG_EPS_F); // custom_layers.py:55
                o[px * R + q] = n4[px] * rcp_approx(den);                // custom_layers.py:56
            }
        }
        store_elems<T, PX * R>(orow + (int64_t)p * out_sH, o);
        if constexpr (D > 0) {
            if (p % D == 0) {
                if (drow) {
                    float dsv[PX * NDS];
#pragma unroll
                    for (int px = 0; px < PX; ++px)
#pragma unroll
                        for (int qq = 0; qq < NDS; ++qq) dsv[px * NDS + qq] = o[px * R + qq * D];
                    store_elems<T, PX * NDS>(drow + (int64_t)(p / D) * ds_sH, dsv);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Vectorised forward.  units = B*h*(w/PX); unit -> (b, i, jg); contiguous coef.
// ------------------------------------------------------------------------------------------------
template <typename T> struct LpgFwdParams {
    const T *coef;
    T *out;
    T *ds;                 // nullable
    int64_t out_sB, out_sH; // element strides of out (column stride 1)
    int64_t ds_sB, ds_sH;
    uint32_t units;
    FastDiv wg, h;          // units per coarse row, coarse rows per image
};

template <typename T, int R, int PX, int D>
__device__ __forceinline__ void lpg_fwd_unit(const LpgFwdParams<T> &prm, uint32_t unit) {
    constexpr int NDS = D ? R / D : 0;
    uint32_t row, jg, b, i;
    prm.wg.divmod(unit, row, jg);
    prm.h.divmod(row, b, i);

    float c[PX * 3];
    load_elems<T, PX * 3, 4>(prm.coef + (size_t)unit * (PX * 3), c);

    float n1[PX], n2[PX], n3[PX], n4[PX];
#pragma unroll
    for (int px = 0; px < PX; ++px) {
        Angles a;
        decode_angles(c[3 * px], c[3 * px + 1], a);
        n1[px] = a.st * a.cp;   // custom_layers.py:50
        n2[px] = a.st * a.sp;
        n3[px] = a.ct;
        n4[px] = c[3 * px + 2];
    }

    T *orow = prm.out + (int64_t)b * prm.out_sB + (int64_t)(i * R) * prm.out_sH + (size_t)jg * (PX * R);
    T *drow = nullptr;
    if constexpr (D > 0) {
        if (prm.ds) drow = prm.ds + (int64_t)b * prm.ds_sB + (int64_t)(i * NDS) * prm.ds_sH + (size_t)jg * (PX * NDS);
    }
    lpg_expand_store<T, R, PX, D>(n1, n2, n3, n4, orow, prm.out_sH, drow, prm.ds_sH);
}

template <typename T, int R, int PX, int D>
__global__ void __launch_bounds__(256) lpg_fwd_vec_kernel(const __grid_constant__ LpgFwdParams<T> prm) {
    const uint32_t unit = blockIdx.x * blockDim.x + threadIdx.x;
    if (unit < prm.units) lpg_fwd_unit<T, R, PX, D>(prm, unit);
}

// Gather G = g_full + scatter(g_ds) for the r x r patches of PX coarse pixels into registers.
template <typename T, int R, int PX, int D>
__device__ __forceinline__ void lpg_load_patch(const T *grow, int64_t gf_sH, const T *drow, int64_t gd_sH, float (&G)[R][PX * R]) {
    constexpr int NDS = D ? R / D : 0;
    if (grow) {
#pragma unroll
        for (int p = 0; p < R; ++p) load_elems<T, PX * R>(grow + (int64_t)p * gf_sH, G[p]);
    } else {
#pragma unroll
        for (int p = 0; p < R; ++p)
#pragma unroll
            for (int k = 0; k < PX * R; ++k) G[p][k] = 0.0f;
    }
    if constexpr (D > 0) {
        if (drow) {
            float t[NDS][PX * NDS];
#pragma unroll
            for (int pp = 0; pp < NDS; ++pp) load_elems<T, PX * NDS>(drow + (int64_t)pp * gd_sH, t[pp]);
#pragma unroll
            for (int pp = 0; pp < NDS; ++pp)
#pragma unroll
                for (int px = 0; px < PX; ++px)
#pragma unroll
                    for (int qq = 0; qq < NDS; ++qq) G[pp * D][px * R + qq * D] += t[pp][px * NDS + qq];
        }
    }
}

// Fixed-order reduction of one r x r patch into the three coefficient gradients:
// columns inside a row, then rows -- all in one thread, so no shuffles and no atomics are needed.
template <int R, int PX>
__device__ __forceinline__ void lpg_reduce_patch(const float (&G)[R][PX * R], int px, float x0, float x1, float x2, float *gout) {
    using Tab = DirTable<R>;
    Angles a;
    decode_angles(x0, x1, a);
    const float n1 = a.st * a.cp, n2 = a.st * a.sp, n3 = a.ct, n4 = x2;
    float g1 = 0.f, g2 = 0.f, g3 = 0.f, g4 = 0.f;
#pragma unroll
    for (int p = 0; p < R; ++p) {
        const float A = fmaf(Tab::off(p), n1, n3);
        float r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int k = p * R + q;
            const float s = fmaf(Tab::off(q), n2, A);
            const float den = fmaf(Tab::w(k), s, BTSLPG_EPS_F);
            const float inv = rcp_approx(den);
            const float u = G[p][px * R + q] * inv;
            const float tq = u * inv;
            r4 += u;
            r1 = fmaf(tq, Tab::u(k), r1);
            r2 = fmaf(tq, Tab::v(k), r2);
            r3 = fmaf(tq, Tab::w(k), r3);
        }
        g1 += r1; g2 += r2; g3 += r3; g4 += r4;
    }
    const float m = -n4;
    g1 *= m; g2 *= m; g3 *= m;
    const float gph = a.st * (g2 * a.cp - g1 * a.sp);
    const float gth = a.ct * (g1 * a.cp + g2 * a.sp) - g3 * a.st;
    gout[0] = (2.0f * BTSLPG_PI_F) * gph;
    gout[1] = (BTSLPG_PI_F / 3.0f) * gth;
    gout[2] = g4;
}

// ------------------------------------------------------------------------------------------------
// Vectorised backward (SURVEY 8(a) a6 + a9).
//   G = g_full + scatter(g_ds);  u = G/den;  g4 = sum u;  tq = u/den
//   g1 = -n4 sum tq*u_pq ; g2 = -n4 sum tq*v_pq ; g3 = -n4 sum tq*w_pq
//   d/dx0 = 2pi * st*(g2*cp - g1*sp) ; d/dx1 = (pi/3) * (ct*(g1*cp + g2*sp) - g3*st) ; d/dx2 = g4
// ------------------------------------------------------------------------------------------------
template <typename T> struct LpgBwdParams {
    const T *coef;
    const T *g_full;  // nullable
    const T *g_ds;    // nullable
    T *g_coef;
    int64_t gf_sB, gf_sH;
    int64_t gd_sB, gd_sH;
    uint32_t units;
    FastDiv wg, h;
};

template <typename T, int R, int PX, int D>
__device__ __forceinline__ void lpg_bwd_unit(const LpgBwdParams<T> &prm, uint32_t unit) {
    constexpr int NDS = D ? R / D : 0;
    uint32_t row, jg, b, i;
    prm.wg.divmod(unit, row, jg);
    prm.h.divmod(row, b, i);

    // issue every load of the patch first (memory-level parallelism), then compute
    float G[R][PX * R];
    const T *grow = prm.g_full ? prm.g_full + (int64_t)b * prm.gf_sB + (int64_t)(i * R) * prm.gf_sH + (size_t)jg * (PX * R) : nullptr;
    const T *drow = nullptr;
    if constexpr (D > 0) {
        if (prm.g_ds) drow = prm.g_ds + (int64_t)b * prm.gd_sB + (int64_t)(i * NDS) * prm.gd_sH + (size_t)jg * (PX * NDS);
    }
    float c[PX * 3];
    load_elems<T, PX * 3, 4>(prm.coef + (size_t)unit * (PX * 3), c);
    lpg_load_patch<T, R, PX, D>(grow, prm.gf_sH, drow, prm.gd_sH, G);

    float gout[PX * 3];
#pragma unroll
    for (int px = 0; px < PX; ++px) lpg_reduce_patch<R, PX>(G, px, c[3 * px], c[3 * px + 1], c[3 * px + 2], &gout[3 * px]);
    store_elems<T, PX * 3, 4>(prm.g_coef + (size_t)unit * (PX * 3), gout);
}

template <typename T, int R, int PX, int D>
__global__ void __launch_bounds__(256) lpg_bwd_vec_kernel(const __grid_constant__ LpgBwdParams<T> prm) {
    const uint32_t unit = blockIdx.x * blockDim.x + threadIdx.x;
    if (unit < prm.units) lpg_bwd_unit<T, R, PX, D>(prm, unit);
}

// ------------------------------------------------------------------------------------------------
// Several layers in one launch (the three scales of the decoder).  Block ranges are assigned per
// layer; the variant switch is block-uniform.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxMulti = 4;
template <typename T> __host__ __device__ constexpr int px_max(int r) { return 32 / (r * (int)sizeof(T)); }

template <typename T> struct LpgFwdMulti {
    LpgFwdParams<T> layer[kMaxMulti];
    uint32_t block_end[kMaxMulti];  // exclusive prefix of blocks per layer
    int32_t upratio[kMaxMulti];
    int32_t n;
};
template <typename T> struct LpgBwdMulti {
    LpgBwdParams<T> layer[kMaxMulti];
    uint32_t block_end[kMaxMulti];
    int32_t upratio[kMaxMulti];
    int32_t n;
};

template <typename T>
__global__ void __launch_bounds__(256) lpg_fwd_multi_kernel(const __grid_constant__ LpgFwdMulti<T> m) {
    int l = 0;
    uint32_t first = 0;
#pragma unroll
    for (int k = 0; k < kMaxMulti - 1; ++k)
        if (k < m.n - 1 && blockIdx.x >= m.block_end[k]) { l = k + 1; first = m.block_end[k]; }
    const LpgFwdParams<T> &prm = m.layer[l];
    const uint32_t unit = (blockIdx.x - first) * blockDim.x + threadIdx.x;
    if (unit >= prm.units) return;
    switch (m.upratio[l]) {
        case 8: lpg_fwd_unit<T, 8, px_max<T>(8), 4>(prm, unit); break;
        case 4: lpg_fwd_unit<T, 4, px_max<T>(4), 2>(prm, unit); break;
        default: lpg_fwd_unit<T, 2, px_max<T>(2), 0>(prm, unit); break;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) lpg_bwd_multi_kernel(const __grid_constant__ LpgBwdMulti<T> m) {
    int l = 0;
    uint32_t first = 0;
#pragma unroll
    for (int k = 0; k < kMaxMulti - 1; ++k)
        if (k < m.n - 1 && blockIdx.x >= m.block_end[k]) { l = k + 1; first = m.block_end[k]; }
    const LpgBwdParams<T> &prm = m.layer[l];
    const uint32_t unit = (blockIdx.x - first) * blockDim.x + threadIdx.x;
    if (unit >= prm.units) return;
    switch (m.upratio[l]) {
        case 8: lpg_bwd_unit<T, 8, px_max<T>(8), 4>(prm, unit); break;
        case 4: lpg_bwd_unit<T, 4, px_max<T>(4), 2>(prm, unit); break;
        default: lpg_bwd_unit<T, 2, px_max<T>(2), 0>(prm, unit); break;
    }
}

// ------------------------------------------------------------------------------------------------
// Generic kernels: any up-ratio r >= 1, any ds stride d | r, any element strides (e.g. an NHWC
// concat slot whose column stride is the channel count).  One thread per coarse pixel, scalar
// accesses.  Same arithmetic as the vectorised kernels (the direction is recomputed with the
// float32 op sequence of custom_layers.py:38-43), so results are bit-identical between paths.
// ------------------------------------------------------------------------------------------------
template <typename T> struct LpgGenericParams {
    const T *coef;
    int64_t c_sB, c_sH, c_sW, c_sC;
    // forward
    T *out;
    T *ds;
    int64_t o_sB, o_sH, o_sW;
    int64_t d_sB, d_sH, d_sW;
    // backward
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
    float g1 = 0.f, g2 = 0.f, g3 = 0.f, g4 = 0.f;
    for (int p = 0; p < r; ++p) {
        const float ap = dir_offset(p, r);
        const float A = fmaf(ap, n1, n3);
        float r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f;
        for (int q = 0; q < r; ++q) {
            const float bq = dir_offset(q, r);
            const float wpq = dir_w(ap, bq);
            float G = 0.f;
            if (prm.g_full) G = load1(prm.g_full + b * prm.o_sB + (i * r + p) * prm.o_sH + (j * r + q) * prm.o_sW);
            if (prm.g_ds && d > 0 && p % d == 0 && q % d == 0)
                G += load1(prm.g_ds + b * prm.d_sB + ((i * r + p) / d) * prm.d_sH + ((j * r + q) / d) * prm.d_sW);
            const float s = fmaf(bq, n2, A);
            const float den = fmaf(wpq, s, BTSLPG_EPS_F);
            const float inv = rcp_approx(den);
            const float u = G * inv;
            const float tq = u * inv;
            r4 += u;
            r1 = fmaf(tq, ap * wpq, r1);
            r2 = fmaf(tq, bq * wpq, r2);
            r3 = fmaf(tq, wpq, r3);
        }
        g1 += r1; g2 += r2; g3 += r3; g4 += r4;
    }
    const float m = -n4;
    g1 *= m; g2 *= m; g3 *= m;
    const float gph = a.st * (g2 * a.cp - g1 * a.sp);
    const float gth = a.ct * (g1 * a.cp + g2 * a.sp) - g3 * a.st;
    T *o = prm.g_coef + b * prm.gc_sB + i * prm.gc_sH + j * prm.gc_sW;
    store1(o, (2.0f * BTSLPG_PI_F) * gph);
    store1(o + prm.gc_sC, (BTSLPG_PI_F / 3.0f) * gth);
    store1(o + 2 * prm.gc_sC, g4);
}

}  // namespace btslpg
