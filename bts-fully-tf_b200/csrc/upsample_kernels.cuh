// upsample_kernels.cuh -- nearest-neighbour x2 up-sampling of NHWC maps (sm_100a): SURVEY 8(f) N1, the
// `layers.UpSampling2D(size=2, interpolation='nearest')` in front of every upconv of the decoder
// (bts_decoder.py:31, :38, :97).  out[b, y, x, :] = in[b, y/2, x/2, :].
//
// Pure data movement: N bytes read, 4N written (forward); 4N read, N written (backward, the sum of the four
// output gradients of each input pixel, added in the fixed order ((0,0)+(0,1)) + ((1,0)+(1,1))).  In the
// reference decoder this is the single largest memory-bound item -- the last one alone writes 2.5 GB at
// B = 32, 480x640 -- and the framework kernel it usually runs on reaches ~0.9 TB/s there.  Here a thread
// owns one 16-byte vector of an input pixel and writes it to the 2 x 2 output pixels (two of them adjacent
// in memory), so every input byte is read once and every store instruction covers whole 128-byte lines.
#pragma once

#include "common.cuh"

namespace btslpg {

constexpr int kUpsampleThreads = 256;

template <typename T> struct UpsampleParams {
    const T *in;       // forward: (B,h,w,C) ; backward: g_out (B,2h,2w,C)
    T *out;            // forward: (B,2h,2w,C) ; backward: g_in (B,h,w,C)
    uint32_t vec_per_px;   // C * sizeof(T) / 16
    uint32_t w;            // input width (pixels)
    uint64_t nvec;         // B*h*w*vec_per_px
    FastDiv div_vpp, div_w;
};

// index of the 16-byte vector of output pixel (row 2*row+dy, col 2*col+dx), `row` counting input rows over the batch
__device__ __forceinline__ uint64_t up_out_index(uint32_t row, uint32_t col, uint32_t v, int dy, int dx, uint32_t w, uint32_t vpp) {
    return (((uint64_t)(2 * row + dy) * (2 * w)) + (2 * col + dx)) * vpp + v;
}

template <typename T> __global__ void __launch_bounds__(kUpsampleThreads) upsample2x_fwd_kernel(const __grid_constant__ UpsampleParams<T> prm) {
    const uint64_t i = (uint64_t)blockIdx.x * kUpsampleThreads + threadIdx.x;
    if (i >= prm.nvec) return;
    // i = (row * w + col) * vpp + v ; rows of all images are stacked (2*row is the output row of the same image)
    uint32_t px, v, row, col;
    prm.div_vpp.divmod((uint32_t)i, px, v);
    prm.div_w.divmod(px, row, col);
    uint32_t wd[4];
    ldg_nc<4>(reinterpret_cast<const uint4 *>(prm.in) + i, wd);
    uint4 *o = reinterpret_cast<uint4 *>(prm.out);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) stg<4>(o + up_out_index(row, col, v, dy, dx, prm.w, prm.vec_per_px), wd);
}

template <typename T> __global__ void __launch_bounds__(kUpsampleThreads) upsample2x_bwd_kernel(const __grid_constant__ UpsampleParams<T> prm) {
    constexpr int N = 16 / (int)sizeof(T);
    const uint64_t i = (uint64_t)blockIdx.x * kUpsampleThreads + threadIdx.x;
    if (i >= prm.nvec) return;
    uint32_t px, v, row, col;
    prm.div_vpp.divmod((uint32_t)i, px, v);
    prm.div_w.divmod(px, row, col);
    float g[4][N];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        load_elems<T, N, 4>(prm.in + up_out_index(row, col, v, k >> 1, k & 1, prm.w, prm.vec_per_px) * N, g[k]);
    float s[N];
#pragma unroll
    for (int e = 0; e < N; ++e) s[e] = (g[0][e] + g[1][e]) + (g[2][e] + g[3][e]);
    store_elems<T, N, 4>(prm.out + i * N, s);
}

// any channel count / alignment: one element per thread
template <typename T> struct UpsampleGenericParams {
    const T *in;
    T *out;
    uint64_t n;        // elements of the SMALL tensor
    uint32_t C, w;
    FastDiv div_c, div_w;
};
template <typename T, bool BWD> __global__ void __launch_bounds__(kUpsampleThreads) upsample2x_generic_kernel(const __grid_constant__ UpsampleGenericParams<T> prm) {
    const uint64_t i = (uint64_t)blockIdx.x * kUpsampleThreads + threadIdx.x;
    if (i >= prm.n) return;
    uint32_t px, c, row, col;
    prm.div_c.divmod((uint32_t)i, px, c);
    prm.div_w.divmod(px, row, col);
    auto big = [&](int dy, int dx) { return (((uint64_t)(2 * row + dy) * (2 * prm.w)) + (2 * col + dx)) * prm.C + c; };
    if constexpr (BWD) {
        store1(prm.out + i, (load1(prm.in + big(0, 0)) + load1(prm.in + big(0, 1))) + (load1(prm.in + big(1, 0)) + load1(prm.in + big(1, 1))));
    } else {
        const float x = load1(prm.in + i);
        store1(prm.out + big(0, 0), x); store1(prm.out + big(0, 1), x); store1(prm.out + big(1, 0), x); store1(prm.out + big(1, 1), x);
    }
}

}  // namespace btslpg
