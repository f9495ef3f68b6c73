// lpg_persist.cuh -- persistent, TMA-staged multi-layer LPG forward.
//
// The one-shot kernels (lpg_kernels.cuh) expose the HBM latency of their coefficient load once per
// thread and hide it with occupancy alone.  Here every warp is persistent and owns a small ring of
// shared-memory stages: the coefficient segment of its next NS-1 items (an item = 32 consecutive
// lane slots of one layer = one contiguous run of 96..768 bytes) is always in flight as a 1-D bulk
// copy with mbarrier completion, so a warp never waits for memory between items and registers are
// not used to hold data in flight.  Items of all layers form one list, item k of warp w being
// w + k*W, so every warp sees the layers in the same proportion (static balance without atomics).
#pragma once

#include "lpg_kernels.cuh"
#include "tma_pipe.cuh"

namespace btslpg {

constexpr int kPersistStages = 4;
constexpr int kPersistStageBytes = 768;   // 32 groups x PX*3 elements, maximum over the variants
constexpr int kPersistWarps = kMultiThreads / 32;

// shared-memory read of N contiguous elements (4- or 8-byte aligned) widened to float
template <typename T, int N> __device__ __forceinline__ void lds_coef(const unsigned char *p, float (&v)[N]) {
    constexpr int NB = N * (int)sizeof(T);
    static_assert(NB % 4 == 0, "whole words");
    constexpr int NW = NB / 4;
    uint32_t w[NW];
    if constexpr (NW % 2 == 0) {
#pragma unroll
        for (int i = 0; i < NW; i += 2) {
            const uint2 u = *reinterpret_cast<const uint2 *>(p + 4 * i);
            w[i] = u.x; w[i + 1] = u.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = *reinterpret_cast<const uint32_t *>(p + 4 * i);
    }
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __uint_as_float(w[i]);
    } else {
#pragma unroll
        for (int i = 0; i < NW; ++i) { v[2 * i] = bf16_lo(w[i]); v[2 * i + 1] = bf16_hi(w[i]); }
    }
}

// per-variant constants of a layer with up-ratio r (forward defaults)
template <typename T> __device__ __forceinline__ void fwd_item_geometry(int r, uint32_t &groups_per_item, uint32_t &bytes_per_group) {
    const int px = px_max<T>(r);
    const int lpp = r / rows_default<T>(r, true);
    groups_per_item = 32 / lpp;
    bytes_per_group = px * 3 * (int)sizeof(T);
}

template <typename T, int R>
__device__ __forceinline__ void lpg_fwd_item(const LpgFwdParams<T> &prm, const float *tab, const unsigned char *stage, bool staged,
                                             uint32_t local, int lane) {
    using V = VecCfg<T, R, true>;
    using S = Split<R, V::ROWS>;
    constexpr int D = R == 2 ? 0 : R / 2;
    const int sub = S::LPP == 1 ? 0 : lane / S::GPW;
    const int gl = S::LPP == 1 ? lane : lane % S::GPW;
    const uint32_t group = local * S::GPW + gl;
    if (group >= prm.groups) return;
    float c[V::PX * 3];
    if (staged) lds_coef<T, V::PX * 3>(stage + gl * (V::PX * 3 * (int)sizeof(T)), c);
    else load_elems<T, V::PX * 3, 4, true>(prm.coef + (size_t)group * (V::PX * 3), c);
    LaneDirs<R, V::ROWS> dir;
    dir.init_from(tab, sub);
    lpg_fwd_compute<T, R, V::PX, V::ROWS, D>(prm, dir, sub, group, c);
}

template <typename T>
__global__ void __launch_bounds__(kMultiThreads, fwd_min_blocks<T>()) lpg_fwd_persist_kernel(const __grid_constant__ LpgFwdMulti<T> m) {
    __shared__ __align__(128) unsigned char ring_all[kPersistWarps][kPersistStages][kPersistStageBytes];
    __shared__ __align__(8) uint64_t bars_all[kPersistWarps][kPersistStages];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char(*ring)[kPersistStageBytes] = ring_all[wid];
    uint64_t *bars = bars_all[wid];
    const uint32_t gw = blockIdx.x * kPersistWarps + wid, W = gridDim.x * kPersistWarps;
    const uint32_t total = m.block_end[m.n - 1];          // block_end holds the item prefix in this kernel

    // bfloat16 r=8 splits its patches over lanes and needs the direction table in shared memory
    const float *tab = nullptr;
    if constexpr (sizeof(T) == 2) {
        bool need = false;
        for (int l = 0; l < m.n; ++l) need |= (m.upratio[l] == 8);
        if (need) tab = stage_dir_table<8>();              // CTA-uniform branch (contains the only CTA barrier)
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kPersistStages; ++k) mbar_init(&bars[k], 1);
        mbar_fence_init();
    }
    __syncwarp();

    // item -> (layer, local item index, staged?, bytes)
    auto locate = [&](uint32_t item, int &l, uint32_t &local) {
        l = 0;
        uint32_t first = 0;
#pragma unroll
        for (int k = 0; k < kMaxMulti - 1; ++k)
            if (k < m.n - 1 && item >= m.block_end[k]) { l = k + 1; first = m.block_end[k]; }
        local = item - first;
    };
    auto item_bytes = [&](int l, uint32_t local, uint32_t &group0) {
        uint32_t gpi, bpg;
        fwd_item_geometry<T>(m.upratio[l], gpi, bpg);
        group0 = local * gpi;
        const uint32_t cnt = min(gpi, m.layer[l].groups - group0);
        return cnt * bpg;
    };
    auto issue = [&](uint32_t k) {                          // lane 0: k-th item of this warp
        const uint32_t item = gw + k * W;
        if (item < total) {
            int l; uint32_t local, group0;
            locate(item, l, local);
            const uint32_t bytes = item_bytes(l, local, group0);
            if ((bytes & 15u) == 0) {                       // whole 16-byte units only (the ragged last item loads directly)
                uint32_t gpi, bpg;
                fwd_item_geometry<T>(m.upratio[l], gpi, bpg);
                uint64_t *bar = &bars[k % kPersistStages];
                mbar_arrive_expect_tx(bar, bytes);
                bulk_g2s(ring[k % kPersistStages], reinterpret_cast<const unsigned char *>(m.layer[l].coef) + (size_t)group0 * bpg, bytes, bar);
            }
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kPersistStages; ++k) issue(k);
    }

    for (uint32_t k = 0, item = gw; item < total; ++k, item += W) {
        int l; uint32_t local, group0;
        locate(item, l, local);
        const bool staged = (item_bytes(l, local, group0) & 15u) == 0;
        const int st = k % kPersistStages;
        if (staged) mbar_wait(&bars[st], (k / kPersistStages) & 1);
        const LpgFwdParams<T> &prm = m.layer[l];
        switch (m.upratio[l]) {
            case 8: lpg_fwd_item<T, 8>(prm, tab, ring[st], staged, local, lane); break;
            case 4: lpg_fwd_item<T, 4>(prm, tab, ring[st], staged, local, lane); break;
            default: lpg_fwd_item<T, 2>(prm, tab, ring[st], staged, local, lane); break;
        }
        __syncwarp();                                       // every lane has taken its coefficients out of the stage
        if (lane == 0) {
            fence_proxy_async();
            issue(k + kPersistStages);
        }
    }
}

}  // namespace btslpg
