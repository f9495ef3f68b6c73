// concat_kernels.cuh -- the decoder's full-resolution concat (bts_decoder.py:98-99) as ONE streaming pass (sm_100a):
//
//     upconv1 = Conv2D(F/16, 3, activation='elu')(upsample1)                              # :98  (the activation)
//     concat1 = Concatenate(axis=3)([upconv1, depth_2x2_scaled, depth_4x4_scaled, depth_8x8_scaled])   # :99
//
// the reference re-copies every input of the concat, after a separate ELU pass over
// the (B,H,W,F/16) conv output.  Here the raw conv output is read once, ELU applied in registers, and the
// F/16 + 3 channel NHWC pixel written once -- with the three LPG planes landing in their channel slots
// (order is load-bearing).  The same kernel covers the conv_block concat [upconv, skip, lpg_ds]
// (bts_decoder.py:42) with activation off.  Backward: d concat -> d conv output (times ELU', taken from
// the saved concat output: y > 0 ? 1 : y + 1), d skip, d planes, again one pass.
//
// Layout problem and answer: an output pixel is CT = CA + CB + NP elements (35 floats = 140 bytes), so
// pixel starts are not 16-byte aligned and neither source maps to the destination with a fixed vector
// shift.  A CTA therefore stages a run of P consecutive pixels in shared memory as the final [P][CT]
// image -- sources are read with flat, aligned 16-byte loads and scattered into it (row stride CT:
// odd strides are bank-conflict-free) -- and then streams the image out with flat, aligned 16-byte
// stores (P is a multiple of 8, so P*CT elements start and end on 16-byte boundaries for both dtypes).
// Algorithmic bytes per pixel: (CA + CB + NP) read + CT written = 2 * CT elements forward;
// backward reads CT (gradient) + CT (saved output, only when the activation is on) and writes CT.
#pragma once

#include "common.cuh"

namespace btslpg {

constexpr int kConcatThreads = 256;
constexpr int kConcatMaxPlanes = 3;
constexpr int kConcatSmemBytes = 24 * 1024;       // per staged image (the backward stages two: 48 KB, no opt-in needed); P is derived from it

template <typename T> struct ConcatParams {
    const T *a;                         // (npix, CA) dense source, activation applied when act != 0
    const T *b;                         // (npix, CB) dense source or NULL
    const T *plane[kConcatMaxPlanes];   // (npix) single-channel sources
    T *out;                             // (npix, CT)
    // backward (same geometry): g_out -> g_a, g_b, g_plane; y = saved forward output (needed when act != 0)
    const T *g_out;
    const T *y;
    T *g_a;
    T *g_b;
    T *g_plane[kConcatMaxPlanes];
    // forward only, optional: per-channel affine applied to source a AFTER the activation -- an inference-mode
    // BatchNormalization folded in: scale = gamma / sqrt(var + eps), shift = beta - mean * scale (bts_decoder.py:33-34, :41)
    const float *scale;
    const float *shift;
    // backward only, optional: training-mode BatchNormalization between the activation and the concat (bnstat_kernels.cuh): the
    // float32 [8][CA] pack {scale, shift, mean, std, 1/gamma, beta, c1 = mean(g), c2 = mean(g * xhat)};
    // g_a = scale * (g - c1 - xhat * c2) * elu'(elu), xhat = (y - beta) / gamma and elu = mean + xhat * std recovered from the output y
    const float *bn;
    uint64_t npix;
    uint32_t ca, cb, np, pad, ct;       // ct = ca + cb + np + pad; the pad channels are written as zeros (and ignored backward)
    uint32_t tile_px;                   // P: pixels per CTA iteration (multiple of 8)
    FastDiv div_ca, div_cb;
    FastDiv div_cpp;                    // chunked forward: 16-byte chunks per output pixel
    int act;                            // 1: ELU(alpha = 1) on source a (Keras activation='elu')
    int vec;                            // CA and CB are multiples of the 16-byte vector width: dense sources use vector accesses
    // Sub-pixel source (sub_w > 0): `a` (and g_a) is the output of a 3x3 convolution run on the LOW-RES input with 4*CA
    // output channels ordered (row parity, column parity, channel) -- algebraically the nearest-x2 up-sampling followed by
    // the 3x3 upconv (bts_decoder.py:97-98), without the up-sampled tensor ever existing.  Its memory is (B, h, w, 2, 2, CA);
    // the concat's pixel (Y, X) reads [row Y/2][col X/2][Y%2][X%2][:], i.e. the pixel shuffle happens in this kernel's addressing.
    uint32_t sub_w;                     // low-res width w (output width 2w); 0 = plain NHWC source
    FastDiv div_w2;                     // 2w
};

// Keras / TF elu: x > 0 ? x : expm1(x).  expm1 for x <= 0 in ~12 branch-free instructions instead of libm's
// ~25 (the pass is 8 bytes per element: every instruction counts): the Taylor series through x^7 on
// [-ln2/2, 0] (truncation < 2e-8 relative) and exp(x) - 1 from MUFU.EX2 below it (result <= -0.29, so the
// SFU's ~1e-7 absolute error is <= 4e-7 relative).  x -> -inf gives -1, NaN propagates.
__device__ __forceinline__ float expm1_nonpos(float x) {
    float p = fmaf(x, 1.0f / 5040.0f, 1.0f / 720.0f);
    p = fmaf(p, x, 1.0f / 120.0f);
    p = fmaf(p, x, 1.0f / 24.0f);
    p = fmaf(p, x, 1.0f / 6.0f);
    p = fmaf(p, x, 0.5f);
    p = fmaf(p, x, 1.0f);
    p = p * x;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.442695040888963407f));
    return x < -0.34657359f ? e - 1.0f : p;
}
__device__ __forceinline__ float elu_fwd(float x) { return x > 0.0f ? x : expm1_nonpos(x); }
// derivative from the OUTPUT y = elu(x): 1 for y > 0, exp(x) = y + 1 otherwise
__device__ __forceinline__ float elu_grad_from_output(float y) { return y > 0.0f ? 1.0f : y + 1.0f; }

template <typename T> __device__ __forceinline__ float smem_get(const T *s, uint32_t i);
template <> __device__ __forceinline__ float smem_get<float>(const float *s, uint32_t i) { return s[i]; }
template <> __device__ __forceinline__ float smem_get<__nv_bfloat16>(const __nv_bfloat16 *s, uint32_t i) { return __bfloat162float(s[i]); }
template <typename T> __device__ __forceinline__ void smem_put(T *s, uint32_t i, float v);
template <> __device__ __forceinline__ void smem_put<float>(float *s, uint32_t i, float v) { s[i] = v; }
template <> __device__ __forceinline__ void smem_put<__nv_bfloat16>(__nv_bfloat16 *s, uint32_t i, float v) { s[i] = __float2bfloat16_rn(v); }

// flat copy global -> shared of n elements (n and both addresses 16-byte granular)
template <typename T> __device__ __forceinline__ void flat_g2s(const T *g, T *s, uint32_t n) {
    constexpr int V = 16 / (int)sizeof(T);
    for (uint32_t i = threadIdx.x * V; i < n; i += kConcatThreads * V) {
        uint32_t w[4];
        ldg_nc<4>(g + i, w);
        *reinterpret_cast<uint4 *>(s + i) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
// the same copy as asynchronous LDGSTS: no registers, every 16-byte piece of the tile in flight at once (the register
// form above compiles to one load in flight per thread: LDG -> STS -> LDG ...).  Finish with flat_g2s_wait().
template <typename T> __device__ __forceinline__ void flat_g2s_async(const T *g, T *s, uint32_t n) {
    constexpr int V = 16 / (int)sizeof(T);
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(s);
    for (uint32_t i = threadIdx.x * V; i < n; i += kConcatThreads * V)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + i * (uint32_t)sizeof(T)), "l"(g + i) : "memory");
}
__device__ __forceinline__ void flat_g2s_wait() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
template <typename T> __device__ __forceinline__ void flat_s2g(const T *s, T *g, uint32_t n) {
    constexpr int V = 16 / (int)sizeof(T);
    for (uint32_t i = threadIdx.x * V; i < n; i += kConcatThreads * V) {
        const uint4 v = *reinterpret_cast<const uint4 *>(s + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        stg<4>(g + i, w);
    }
}

// scatter a dense source tile (npx pixels x C channels, contiguous) into image rows at channel offset c0
template <typename T, bool ACT> __device__ __forceinline__ void scatter_dense(const T *src, T *img, uint32_t npx, uint32_t C, const FastDiv &divC,
                                                                              uint32_t ct, uint32_t c0, const float *aff = nullptr) {
    constexpr int V = 16 / (int)sizeof(T);
    const uint32_t n = npx * C;                        // a multiple of V, at least V
    // two 16-byte loads in flight per thread.  The second index is CLAMPED (always a legal address; its value is used only
    // in range) rather than the load made conditional: under the kernel's 32-register cap the conditional form spills the
    // loaded vector right behind the load (the warp then waits for it before the first load is even issued).
    for (uint32_t i = threadIdx.x * V; i < n; i += 2 * kConcatThreads * V) {
        const uint32_t j = i + kConcatThreads * V;
        float v[2][V];
        load_elems<T, V, 4>(src + i, v[0]);
        load_elems<T, V, 4>(src + min(j, n - V), v[1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t iu = u ? j : i;
            if (iu >= n) break;
            uint32_t p, c;
            divC.divmod(iu, p, c);                     // C is a multiple of V: the V elements belong to one pixel
#pragma unroll
            for (int e = 0; e < V; ++e) {
                float x = ACT ? elu_fwd(v[u][e]) : v[u][e];
                if (aff) x = fmaf(x, aff[c + e], aff[C + c + e]);      // shared-memory copy of (scale, shift)
                smem_put<T>(img, p * ct + c0 + c + e, x);
            }
        }
    }
}

// element offset of channel 0 of output pixel q (flat index over B * 2h * 2w) in the un-shuffled (B,h,w,2,2,CA) tensor
template <typename T> __device__ __forceinline__ size_t subpixel_offset(const ConcatParams<T> &prm, uint64_t q) {
    uint32_t rowY, X;
    prm.div_w2.divmod((uint32_t)q, rowY, X);
    return (((size_t)(rowY >> 1) * prm.sub_w + (X >> 1)) * 4 + ((rowY & 1) << 1 | (X & 1))) * prm.ca;
}

// 8 CTAs per SM (a 32-register cap, no spills): the affine / sub-pixel / pad additions had pushed the kernel to 40 registers,
// i.e. 6 resident CTAs, and the plain concat1 pass from 565 us to 640 us (B = 32, 480x640)
template <typename T> __global__ void __launch_bounds__(kConcatThreads, 8) concat_fwd_kernel(const __grid_constant__ ConcatParams<T> prm) {
    extern __shared__ __align__(16) unsigned char concat_smem[];
    T *img = reinterpret_cast<T *>(concat_smem);
    const uint32_t P = prm.tile_px, ct = prm.ct;
    // (scale, shift) of the folded BatchNormalization live behind the staged image
    float *aff = nullptr;
    if (prm.scale) {
        aff = reinterpret_cast<float *>(concat_smem + (((size_t)P * ct * sizeof(T) + 15) / 16) * 16);
        for (uint32_t c = threadIdx.x; c < prm.ca; c += kConcatThreads) {
            aff[c] = prm.scale[c];
            aff[prm.ca + c] = prm.shift[c];
        }
        __syncthreads();
    }
    const uint64_t ntiles = (prm.npix + P - 1) / P;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t p0 = t * P;
        const uint32_t npx = (uint32_t)min((uint64_t)P, prm.npix - p0);
        const bool full_out = (npx % 8) == 0;          // vector paths need 16-byte granular runs
        const bool full = full_out && prm.vec;
        if (prm.sub_w) {                               // source a in the un-shuffled sub-pixel layout (CA % V == 0 checked on the host)
            constexpr int V = 16 / (int)sizeof(T);
            for (uint32_t i = threadIdx.x * V; i < npx * prm.ca; i += kConcatThreads * V) {
                uint32_t p, c;
                prm.div_ca.divmod(i, p, c);
                float v[V];
                load_elems<T, V, 4>(prm.a + subpixel_offset(prm, p0 + p) + c, v);
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    float x = prm.act ? elu_fwd(v[e]) : v[e];
                    if (aff) x = fmaf(x, aff[c + e], aff[prm.ca + c + e]);
                    smem_put<T>(img, p * ct + c + e, x);
                }
            }
            if (prm.b) {
                if (full) scatter_dense<T, false>(prm.b + p0 * prm.cb, img, npx, prm.cb, prm.div_cb, ct, prm.ca);
                else
                    for (uint32_t i = threadIdx.x; i < npx * prm.cb; i += kConcatThreads) {
                        uint32_t p, c;
                        prm.div_cb.divmod(i, p, c);
                        smem_put<T>(img, p * ct + prm.ca + c, load1(prm.b + p0 * prm.cb + i));
                    }
            }
        } else if (full) {
            if (prm.act) scatter_dense<T, true>(prm.a + p0 * prm.ca, img, npx, prm.ca, prm.div_ca, ct, 0, aff);
            else scatter_dense<T, false>(prm.a + p0 * prm.ca, img, npx, prm.ca, prm.div_ca, ct, 0, aff);
            if (prm.b) scatter_dense<T, false>(prm.b + p0 * prm.cb, img, npx, prm.cb, prm.div_cb, ct, prm.ca);
        } else {
            for (uint32_t i = threadIdx.x; i < npx * prm.ca; i += kConcatThreads) {
                uint32_t p, c;
                prm.div_ca.divmod(i, p, c);
                float v = load1(prm.a + p0 * prm.ca + i);
                if (prm.act) v = elu_fwd(v);
                if (aff) v = fmaf(v, aff[c], aff[prm.ca + c]);
                smem_put<T>(img, p * ct + c, v);
            }
            if (prm.b)
                for (uint32_t i = threadIdx.x; i < npx * prm.cb; i += kConcatThreads) {
                    uint32_t p, c;
                    prm.div_cb.divmod(i, p, c);
                    smem_put<T>(img, p * ct + prm.ca + c, load1(prm.b + p0 * prm.cb + i));
                }
        }
        for (uint32_t k = 0; k < prm.np; ++k)
            for (uint32_t p = threadIdx.x; p < npx; p += kConcatThreads) smem_put<T>(img, p * ct + prm.ca + prm.cb + k, load1(prm.plane[k] + p0 + p));
        for (uint32_t i = threadIdx.x; i < npx * prm.pad; i += kConcatThreads)      // zero channels that align CT (cuDNN would pad otherwise)
            smem_put<T>(img, (i / prm.pad) * ct + prm.ca + prm.cb + prm.np + (i % prm.pad), 0.0f);
        __syncthreads();
        if (full_out) {
            flat_s2g<T>(img, prm.out + p0 * ct, npx * ct);
        } else {
            for (uint32_t i = threadIdx.x; i < npx * ct; i += kConcatThreads) prm.out[p0 * ct + i] = img[i];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Chunked forward: when CA, CB and n_planes + pad are whole 16-byte chunks (every concat of the decoder: 512+384, 256+192,
// 128+96+1+3, 64+96+1+3, 32+0+3+1) each 16-byte chunk of the output comes from exactly ONE source, and the output is
// simply a linear stream of chunks.  One thread per chunk: a vector load from a (sub-pixel addressing, ELU, folded
// BatchNormalization) or from b, or three scalar plane loads plus zeros for the last chunk of the pixel, then one vector
// store -- no shared-memory staging, no scalar stores, U chunks per thread in flight.  EXPERIMENT (tuning key 10 = 1): on the
// decoder's five concats it measured slower than the staged kernel above (1954 us vs 1709 us per inference step at B = 32,
// 480x640), which stays the default.
// ------------------------------------------------------------------------------------------------
constexpr int kConcatChunkTilePx = 1024;     // pixels per work item: chunk indices inside an item fit FastDiv's 31 bits
constexpr int kConcatChunkUnroll = 4;

template <typename T, bool ACT> __global__ void __launch_bounds__(kConcatThreads) concat_fwd_chunk_kernel(const __grid_constant__ ConcatParams<T> prm) {
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int U = kConcatChunkUnroll;
    const uint32_t cpp = prm.ct / V, ka = prm.ca / V, kab = (prm.ca + prm.cb) / V;      // chunks per pixel; first chunk of b / of the planes
    const FastDiv div_cpp = prm.div_cpp;
    const uint64_t ntiles = (prm.npix + kConcatChunkTilePx - 1) / kConcatChunkTilePx;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t p0 = t * kConcatChunkTilePx;
        const uint32_t npx = (uint32_t)min((uint64_t)kConcatChunkTilePx, prm.npix - p0);
        const uint32_t n = npx * cpp;
        T *out = prm.out + p0 * prm.ct;
        for (uint32_t base = threadIdx.x; base < n; base += kConcatThreads * U) {
            float v[U][V];
            uint32_t kk[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t i = base + u * kConcatThreads;
                kk[u] = 0xffffffffu;
                if (i < n) {
                    uint32_t p, k;
                    div_cpp.divmod(i, p, k);
                    kk[u] = k;
                    if (k < ka) {
                        const size_t off = prm.sub_w ? subpixel_offset(prm, p0 + p) : (p0 + p) * prm.ca;
                        load_elems<T, V, 4>(prm.a + off + k * V, v[u]);
                    } else if (k < kab) {
                        load_elems<T, V, 4>(prm.b + (p0 + p) * prm.cb + (k - ka) * V, v[u]);
                    } else {
#pragma unroll
                        for (int e = 0; e < V; ++e) v[u][e] = 0.0f;                     // the pad channels
#pragma unroll
                        for (int e = 0; e < kConcatMaxPlanes; ++e)
                            if (e < (int)prm.np) v[u][e] = load1(prm.plane[e] + p0 + p);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t i = base + u * kConcatThreads;
                if (kk[u] == 0xffffffffu) continue;
                if (kk[u] < ka) {
                    if constexpr (ACT) {
#pragma unroll
                        for (int e = 0; e < V; ++e) v[u][e] = elu_fwd(v[u][e]);
                    }
                    if (prm.scale) {
                        const uint32_t c = kk[u] * V;
#pragma unroll
                        for (int e = 0; e < V; e += 4) {
                            const float4 sc = __ldg(reinterpret_cast<const float4 *>(prm.scale + c + e));
                            const float4 sh = __ldg(reinterpret_cast<const float4 *>(prm.shift + c + e));
                            v[u][e] = fmaf(v[u][e], sc.x, sh.x);
                            v[u][e + 1] = fmaf(v[u][e + 1], sc.y, sh.y);
                            v[u][e + 2] = fmaf(v[u][e + 2], sc.z, sh.z);
                            v[u][e + 3] = fmaf(v[u][e + 3], sc.w, sh.w);
                        }
                    }
                }
                store_elems<T, V, 4>(out + (size_t)i * V, v[u]);
            }
        }
    }
}

// gradient of source a for one element of channel c: plain slice, ELU' from the output, or the training-mode BatchNorm form
template <typename T> __device__ __forceinline__ float concat_ga(const ConcatParams<T> &prm, float g, float y, uint32_t c) {
    if (prm.bn) {
        const uint32_t ca = prm.ca;
        const float xhat = (y - __ldg(prm.bn + 5 * ca + c)) * __ldg(prm.bn + 4 * ca + c);
        const float elu = fmaf(xhat, __ldg(prm.bn + 3 * ca + c), __ldg(prm.bn + 2 * ca + c));
        const float d = __ldg(prm.bn + c) * (g - __ldg(prm.bn + 6 * ca + c) - xhat * __ldg(prm.bn + 7 * ca + c));
        return prm.act ? d * elu_grad_from_output(elu) : d;
    }
    return prm.act ? g * elu_grad_from_output(y) : g;
}

template <typename T> __global__ void __launch_bounds__(kConcatThreads) concat_bwd_kernel(const __grid_constant__ ConcatParams<T> prm) {
    extern __shared__ __align__(16) unsigned char concat_smem[];
    const uint32_t P = prm.tile_px, ct = prm.ct;
    T *gimg = reinterpret_cast<T *>(concat_smem);
    T *yimg = gimg + (size_t)P * ct;                    // only used when the activation is on
    constexpr int V = 16 / (int)sizeof(T);
    const uint64_t ntiles = (prm.npix + P - 1) / P;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t p0 = t * P;
        const uint32_t npx = (uint32_t)min((uint64_t)P, prm.npix - p0);
        const bool full_in = (npx % 8) == 0;
        const bool full = full_in && prm.vec;
        if (full_in) {
            flat_g2s_async<T>(prm.g_out + p0 * ct, gimg, npx * ct);
            if (prm.act || prm.bn) flat_g2s_async<T>(prm.y + p0 * ct, yimg, npx * ct);
            flat_g2s_wait();
        } else {
            for (uint32_t i = threadIdx.x; i < npx * ct; i += kConcatThreads) {
                gimg[i] = prm.g_out[p0 * ct + i];
                if (prm.act || prm.bn) yimg[i] = prm.y[p0 * ct + i];
            }
        }
        __syncthreads();
        if (prm.g_a && prm.sub_w) {                       // gradient in the un-shuffled sub-pixel layout
            const uint32_t n = npx * prm.ca;
            for (uint32_t i = threadIdx.x * V; i < n; i += kConcatThreads * V) {
                uint32_t p, c;
                prm.div_ca.divmod(i, p, c);
                float v[V];
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const float g = smem_get<T>(gimg, p * ct + c + e);
                    v[e] = prm.act ? g * elu_grad_from_output(smem_get<T>(yimg, p * ct + c + e)) : g;
                }
                store_elems<T, V, 4>(prm.g_a + subpixel_offset(prm, p0 + p) + c, v);
            }
        } else if (prm.g_a) {
            if (full) {
                const uint32_t n = npx * prm.ca;
                for (uint32_t i = threadIdx.x * V; i < n; i += kConcatThreads * V) {
                    uint32_t p, c;
                    prm.div_ca.divmod(i, p, c);
                    float v[V];
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const float g = smem_get<T>(gimg, p * ct + c + e);
                        v[e] = concat_ga<T>(prm, g, (prm.act || prm.bn) ? smem_get<T>(yimg, p * ct + c + e) : 0.0f, c + e);
                    }
                    store_elems<T, V, 4>(prm.g_a + p0 * prm.ca + i, v);
                }
            } else {
                for (uint32_t i = threadIdx.x; i < npx * prm.ca; i += kConcatThreads) {
                    uint32_t p, c;
                    prm.div_ca.divmod(i, p, c);
                    const float g = smem_get<T>(gimg, p * ct + c);
                    store1(prm.g_a + p0 * prm.ca + i, concat_ga<T>(prm, g, (prm.act || prm.bn) ? smem_get<T>(yimg, p * ct + c) : 0.0f, c));
                }
            }
        }
        if (prm.g_b) {
            if (full) {
                const uint32_t n = npx * prm.cb;
                for (uint32_t i = threadIdx.x * V; i < n; i += kConcatThreads * V) {
                    uint32_t p, c;
                    prm.div_cb.divmod(i, p, c);
                    float v[V];
#pragma unroll
                    for (int e = 0; e < V; ++e) v[e] = smem_get<T>(gimg, p * ct + prm.ca + c + e);
                    store_elems<T, V, 4>(prm.g_b + p0 * prm.cb + i, v);
                }
            } else {
                for (uint32_t i = threadIdx.x; i < npx * prm.cb; i += kConcatThreads) {
                    uint32_t p, c;
                    prm.div_cb.divmod(i, p, c);
                    store1(prm.g_b + p0 * prm.cb + i, smem_get<T>(gimg, p * ct + prm.ca + c));
                }
            }
        }
        for (uint32_t k = 0; k < prm.np; ++k)
            if (prm.g_plane[k])
                for (uint32_t p = threadIdx.x; p < npx; p += kConcatThreads) store1(prm.g_plane[k] + p0 + p, smem_get<T>(gimg, p * ct + prm.ca + prm.cb + k));
        __syncthreads();
    }
}

}  // namespace btslpg
