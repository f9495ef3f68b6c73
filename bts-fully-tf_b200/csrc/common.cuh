// common.cuh -- small device utilities shared by the LPG kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace btslpg {

// python `pi` meeting a float32 tensor (reference custom_layers.py:49) and K.epsilon() (:55)
#define BTSLPG_PI_F 3.14159265358979323846f
#define BTSLPG_EPS_F 1e-7f

// ------------------------------------------------------------------------------------------------
// Division of a 31-bit unsigned by a run-time constant in three instructions, no special cases
// (Granlund-Montgomery "add" form): l = ceil(log2 d), m = floor(2^32 (2^l - d) / d) + 1,
// q = (umulhi(n, m) + n) >> l.  umulhi(n, m) < n < 2^31, so the sum cannot overflow.  d = 1 gives
// m = 1, l = 0 -> q = n.
// ------------------------------------------------------------------------------------------------
struct FastDiv {
    uint32_t d, mul, shr;
    FastDiv() : d(1), mul(1), shr(0) {}
    explicit FastDiv(uint32_t div) : d(div ? div : 1), mul(1), shr(0) {
        uint32_t l = 0;
        while ((1ull << l) < d) ++l;                     // ceil(log2(d))
        mul = (uint32_t)((((1ull << l) - d) << 32) / d) + 1;
        shr = l;
    }
    __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
        return (__umulhi(n, mul) + n) >> shr;
#else
        return (uint32_t)(((((uint64_t)n * mul) >> 32) + n) >> shr);
#endif
    }
    __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t &q, uint32_t &rem) const {
        q = div(n);
        rem = n - q * d;
    }
};

// ------------------------------------------------------------------------------------------------
// Element <-> float32 conversion through raw 32-bit words (two bf16 per word, little endian).
// ------------------------------------------------------------------------------------------------
template <typename T> struct ElemTraits;
template <> struct ElemTraits<float> {
    static constexpr int kBytes = 4;
    static constexpr const char *kName = "f32";
};
template <> struct ElemTraits<__nv_bfloat16> {
    static constexpr int kBytes = 2;
    static constexpr const char *kName = "bf16";
};

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    // round-to-nearest-even, hi in the upper half: cvt.rn.bf16x2.f32 d, a(hi), b(lo)
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// ------------------------------------------------------------------------------------------------
// Global memory access in 32-bit words with explicit widths.  VW = words per instruction
// (1, 2, 4, 8 -> LDG/STG .32/.64/.128/.256; the 256-bit forms are new on sm_100).
// Loads use the read-only path without L1 allocation (every input byte is read exactly once);
// stores are plain write-back so that a consumer kernel can still find them in the 126 MB L2.
// ------------------------------------------------------------------------------------------------
template <int VW> __device__ __forceinline__ void ldg_nc(const void *p, uint32_t *w);
template <> __device__ __forceinline__ void ldg_nc<1>(const void *p, uint32_t *w) {
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(w[0]) : "l"(p));
}
template <> __device__ __forceinline__ void ldg_nc<2>(const void *p, uint32_t *w) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(p));
}
template <> __device__ __forceinline__ void ldg_nc<4>(const void *p, uint32_t *w) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                 : "l"(p));
}
template <> __device__ __forceinline__ void ldg_nc<8>(const void *p, uint32_t *w) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
}

// Same, but allocating in L1: for the 12-byte coefficient pixels, whose three words are fetched by
// separate instructions (and, when a patch is split over lanes, by several lanes) -- the first touch
// brings the sector, the others hit.
template <int VW> __device__ __forceinline__ void ldg_nc_l1(const void *p, uint32_t *w);
template <> __device__ __forceinline__ void ldg_nc_l1<1>(const void *p, uint32_t *w) {
    asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(w[0]) : "l"(p));
}
template <> __device__ __forceinline__ void ldg_nc_l1<2>(const void *p, uint32_t *w) {
    asm volatile("ld.global.nc.v2.b32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(p));
}
template <> __device__ __forceinline__ void ldg_nc_l1<4>(const void *p, uint32_t *w) {
    asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
}

template <int VW> __device__ __forceinline__ void stg(void *p, const uint32_t *w);
template <> __device__ __forceinline__ void stg<1>(void *p, const uint32_t *w) {
    asm volatile("st.global.b32 [%0], %1;" ::"l"(p), "r"(w[0]) : "memory");
}
template <> __device__ __forceinline__ void stg<2>(void *p, const uint32_t *w) {
    asm volatile("st.global.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(w[0]), "r"(w[1]) : "memory");
}
template <> __device__ __forceinline__ void stg<4>(void *p, const uint32_t *w) {
    asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                 : "memory");
}
template <> __device__ __forceinline__ void stg<8>(void *p, const uint32_t *w) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                 "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}

__host__ __device__ constexpr int cmin(int a, int b) { return a < b ? a : b; }
// widest power-of-two word count (<= cap) that divides nwords
__host__ __device__ constexpr int vec_words(int nwords, int cap) {
    int v = 1;
    while (v * 2 <= cap && nwords % (v * 2) == 0) v *= 2;
    return v;
}

// Load N contiguous elements of T (N*sizeof(T) a multiple of 4 bytes; address aligned to the chosen
// vector width, which the host guarantees) and widen to float.
template <typename T, int N, int CAPW = 8, bool L1 = false> __device__ __forceinline__ void load_elems(const T *p, float (&v)[N]) {
    constexpr int NB = N * (int)sizeof(T);
    static_assert(NB % 4 == 0, "load_elems needs whole 32-bit words");
    constexpr int NW = NB / 4;
    constexpr int VW = vec_words(NW, L1 ? cmin(CAPW, 4) : CAPW);
    uint32_t w[NW];
#pragma unroll
    for (int i = 0; i < NW; i += VW) {
        if constexpr (L1) ldg_nc_l1<VW>(reinterpret_cast<const uint32_t *>(p) + i, w + i);
        else ldg_nc<VW>(reinterpret_cast<const uint32_t *>(p) + i, w + i);
    }
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __uint_as_float(w[i]);
    } else {
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            v[2 * i] = bf16_lo(w[i]);
            v[2 * i + 1] = bf16_hi(w[i]);
        }
    }
}

template <typename T, int N, int CAPW = 8> __device__ __forceinline__ void store_elems(T *p, const float (&v)[N]) {
    constexpr int NB = N * (int)sizeof(T);
    static_assert(NB % 4 == 0, "store_elems needs whole 32-bit words");
    constexpr int NW = NB / 4;
    constexpr int VW = vec_words(NW, CAPW);
    uint32_t w[NW];
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < N; ++i) w[i] = __float_as_uint(v[i]);
    } else {
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    }
#pragma unroll
    for (int i = 0; i < NW; i += VW) stg<VW>(reinterpret_cast<uint32_t *>(p) + i, w + i);
}

// scalar element access for the generic (any-stride) kernels
__device__ __forceinline__ float load1(const float *p) { return __ldg(p); }
__device__ __forceinline__ float load1(const __nv_bfloat16 *p) {
    return __uint_as_float(((uint32_t) * reinterpret_cast<const unsigned short *>(p)) << 16);
}
__device__ __forceinline__ void store1(float *p, float v) { *p = v; }
__device__ __forceinline__ void store1(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// 1/x from the SFU (one MUFU.RCP, <= 1 ulp) -- keeps IEEE behaviour at the points that matter for the
// layer: rcp(+-0) = +-inf, rcp(+-inf) = +-0, NaN propagates.  The reference does no clamping of the
// denominator (custom_layers.py:55-56) and neither do we.  The .ftz form is used because the plain
// form costs ~5 extra instructions per call for subnormal operands, which the denominator
// den = w*s + 1e-7 cannot be: it is either exactly 0 or at least one ulp of 1e-7 (~7e-15) away from it.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------------------------------------
// Packed float32 pairs: Blackwell (sm_100) issues two IEEE fp32 FMAs / multiplies / adds in ONE
// instruction on a 64-bit register pair (PTX fma.rn.f32x2, SASS FFMA2).  These kernels are limited by
// instruction issue before they are limited by HBM, so everything that is done for two adjacent
// pixels (or for the two angles of a coefficient) is done as a pair.  Per-lane results are bit-identical
// to the scalar fmaf / * / + they replace.
// ------------------------------------------------------------------------------------------------
struct F2 {
    unsigned long long v;
};
__device__ __forceinline__ F2 f2(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ F2 f2(float x) { return f2(x, x); }
__device__ __forceinline__ F2 f2(float2 x) { return f2(x.x, x.y); }
__device__ __forceinline__ void unpack(F2 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ float lo(F2 a) { float l, h; unpack(a, l, h); return l; }
__device__ __forceinline__ float hi(F2 a) { float l, h; unpack(a, l, h); return h; }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
    F2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 rcp2(F2 a) {      // two MUFU.RCP; the pair stays in adjacent registers
    float l, h;
    unpack(a, l, h);
    return f2(rcp_approx(l), rcp_approx(h));
}

// an int as a type (tag dispatch of compile-time variants)
template <int V> struct IntC {
    static constexpr int value = V;
};

}  // namespace btslpg
