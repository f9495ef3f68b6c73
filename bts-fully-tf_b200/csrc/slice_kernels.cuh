// slice_kernels.cuh -- per-channel affine + activation copy between channel slices of NHWC tensors (sm_100a).
//
// SURVEY 8(f) N3, the DenseASPP block of the decoder (bts_decoder.py:46-54, :61-76): the reference
//   * re-copies a growing feature map five times (Concatenate [iconv4, daspp_3], [.., daspp_6], ...: 384 -> 896 channels),
//   * then reads and writes it again for BatchNormalization and once more for ReLU before each 1x1 conv.
// With one (B,h,w,896) buffer that the blocks append to, all of that is ONE pass per block: read the first Ck
// channels of every pixel (a strided slice), apply the folded inference BatchNormalization and the ReLU, and
// write the contiguous (B,h,w,Ck) input of the 1x1 conv.  The same kernel writes a block's 128 new channels
// into their slot (plain strided copy), puts ELU(iconv4) into the first 256 channels, and finally turns those into
// iconv4_bn in place for concat4_daspp (bts_decoder.py:75).
//
//   dst[p, c] = act(src[p, c] * scale[c] + shift[c])      p over pixels (any pixel stride on both sides), c < C
// act: 0 none, 1 ELU, 2 ReLU; scale/shift optional.  Algorithmic bytes per pixel: 2 * C * sizeof(T).
#pragma once

#include "common.cuh"
#include "concat_kernels.cuh"   // elu_fwd

namespace btslpg {

constexpr int kSliceThreads = 256;

template <typename T> struct SliceParams {
    const T *src;
    T *dst;
    const float *scale;   // nullable (together with shift)
    const float *shift;
    uint64_t n;           // vectors (vec path) or elements (scalar path)
    uint32_t C;           // channels
    uint32_t per_px;      // vectors (or elements) per pixel
    int64_t s_src, s_dst; // pixel strides in elements
    FastDiv div_pp;
    int act;
    // sub-grid ("space to batch") addressing of ONE side: 0 none, 1 src, 2 dst.  That side is a (B*s*s, H/s, W/s, C) tensor whose image
    // (b*s + i)*s + j is the sub-grid [i::s, j::s] of image b of the other side's (B,H,W,C) pixels (decoder._s2b).
    int split;
    uint32_t s, H, W;
    FastDiv div_hw, div_w, div_s;
};

// pixel p = (b, y, x) of the (B,H,W) side -> its pixel index on the sub-grid side
template <typename T> __device__ __forceinline__ uint32_t slice_split_pixel(const SliceParams<T> &prm, uint32_t p) {
    uint32_t b, rem, y, x, ys, yi, xs, xi;
    prm.div_hw.divmod(p, b, rem);
    prm.div_w.divmod(rem, y, x);
    prm.div_s.divmod(y, ys, yi);
    prm.div_s.divmod(x, xs, xi);
    const uint32_t hs = prm.H / prm.s, ws = prm.W / prm.s;
    return (((b * prm.s + yi) * prm.s + xi) * hs + ys) * ws + xs;
}

__device__ __forceinline__ float slice_apply(float x, float sc, float sh, int act) {
    x = fmaf(x, sc, sh);
    return act == 1 ? elu_fwd(x) : act == 2 ? fmaxf(x, 0.0f) : x;
}

template <typename T> __global__ void __launch_bounds__(kSliceThreads) slice_affine_act_vec_kernel(const __grid_constant__ SliceParams<T> prm) {
    constexpr int N = 16 / (int)sizeof(T);
    const uint64_t i = (uint64_t)blockIdx.x * kSliceThreads + threadIdx.x;
    if (i >= prm.n) return;
    uint32_t p, v;
    prm.div_pp.divmod((uint32_t)i, p, v);
    float x[N];
    uint32_t ps = p, pd = p;
    if (prm.split == 1) ps = slice_split_pixel(prm, p);
    else if (prm.split == 2) pd = slice_split_pixel(prm, p);
    load_elems<T, N, 4>(prm.src + (int64_t)ps * prm.s_src + v * N, x);
    if (prm.scale) {
#pragma unroll
        for (int e = 0; e < N; ++e) x[e] = slice_apply(x[e], __ldg(prm.scale + v * N + e), __ldg(prm.shift + v * N + e), prm.act);
    } else {
#pragma unroll
        for (int e = 0; e < N; ++e) x[e] = slice_apply(x[e], 1.0f, 0.0f, prm.act);
    }
    store_elems<T, N, 4>(prm.dst + (int64_t)pd * prm.s_dst + v * N, x);
}

template <typename T> __global__ void __launch_bounds__(kSliceThreads) slice_affine_act_scalar_kernel(const __grid_constant__ SliceParams<T> prm) {
    const uint64_t i = (uint64_t)blockIdx.x * kSliceThreads + threadIdx.x;
    if (i >= prm.n) return;
    uint32_t p, c;
    prm.div_pp.divmod((uint32_t)i, p, c);
    const float sc = prm.scale ? __ldg(prm.scale + c) : 1.0f, sh = prm.scale ? __ldg(prm.shift + c) : 0.0f;
    uint32_t ps = p, pd = p;
    if (prm.split == 1) ps = slice_split_pixel(prm, p);
    else if (prm.split == 2) pd = slice_split_pixel(prm, p);
    store1(prm.dst + (int64_t)pd * prm.s_dst + c, slice_apply(load1(prm.src + (int64_t)ps * prm.s_src + c), sc, sh, prm.act));
}

}  // namespace btslpg
