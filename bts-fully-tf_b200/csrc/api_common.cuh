// api_common.cuh -- helpers shared by the translation units of libbtslpg.so (one .cu per kernel family, so that
// nvcc compiles them in parallel): thread-local error / kernel-name text, tensor views, shape checks.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/btslpg.h"

namespace btslpg_api {

extern thread_local char tl_error[512];
extern thread_local char tl_kernel[128];
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_fwd_threads, g_bwd_threads;
extern std::atomic<int> g_tune_head_impl, g_tune_concat_impl, g_tune_depthconv_impl;

int fail(int code, const char *fmt, ...);

enum DType { kF32 = 0, kBF16 = 1 };

// A tensor viewed as (B, H, W, C) with element strides.
struct View {
    char *ptr = nullptr;
    int dtype = -1;
    int dev = -1;
    int64_t B = 0, H = 0, W = 0, C = 0;
    int64_t sB = 0, sH = 0, sW = 0, sC = 0;
    int esize() const { return dtype == kF32 ? 4 : 2; }
    bool aligned(int bytes) const { return (reinterpret_cast<uintptr_t>(ptr) % bytes) == 0; }
};

inline int parse_common(const BtsTensor *t, const char *name, View &v) {
    if (!t) return fail(BTSLPG_EINVAL, "%s: tensor is NULL", name);
    if (!t->data && t->ndim > 0) {
        int64_t n = 1;
        for (int k = 0; k < t->ndim; ++k) n *= t->shape[k];
        if (n != 0) return fail(BTSLPG_EINVAL, "%s: data pointer is NULL", name);
    }
    if (t->device.device_type != 2 && t->device.device_type != 13)
        return fail(BTSLPG_EDEVICE, "%s: not a CUDA tensor (DLPack device_type %d); host tensors are not accepted -- "
                                    "there is no CPU fallback", name, (int)t->device.device_type);
    if (t->dtype.lanes != 1) return fail(BTSLPG_EDTYPE, "%s: dtype lanes must be 1", name);
    if (t->dtype.code == 2 && t->dtype.bits == 32) v.dtype = kF32;
    else if (t->dtype.code == 4 && t->dtype.bits == 16) v.dtype = kBF16;
    else return fail(BTSLPG_EDTYPE, "%s: dtype (code %d, bits %d) is not float32 or bfloat16", name, (int)t->dtype.code, (int)t->dtype.bits);
    v.dev = t->device.device_id;
    v.ptr = static_cast<char *>(t->data) + t->byte_offset;
    if (!t->shape) return fail(BTSLPG_ESHAPE, "%s: shape is NULL", name);
    return 0;
}

inline void strides_of(const BtsTensor *t, int64_t *s) {
    if (t->strides) {
        for (int k = 0; k < t->ndim; ++k) s[k] = t->strides[k];
    } else {
        int64_t acc = 1;
        for (int k = t->ndim - 1; k >= 0; --k) { s[k] = acc; acc *= t->shape[k]; }
    }
}

// (B,h,w,C) NHWC tensor
inline int parse_nhwc(const BtsTensor *t, const char *name, View &v) {
    if (int e = parse_common(t, name, v)) return e;
    if (t->ndim != 4) return fail(BTSLPG_ESHAPE, "%s: expected a rank-4 NHWC tensor, got rank %d", name, (int)t->ndim);
    int64_t s[4];
    strides_of(t, s);
    v.B = t->shape[0]; v.H = t->shape[1]; v.W = t->shape[2]; v.C = t->shape[3];
    v.sB = s[0]; v.sH = s[1]; v.sW = s[2]; v.sC = s[3];
    if (v.B < 0 || v.H < 0 || v.W < 0 || v.C < 0) return fail(BTSLPG_ESHAPE, "%s: negative extent", name);
    return 0;
}

// single-channel map given as (B,H,W,1) or (B,H,W)
inline int parse_map(const BtsTensor *t, const char *name, View &v) {
    if (int e = parse_common(t, name, v)) return e;
    if (t->ndim != 3 && t->ndim != 4)
        return fail(BTSLPG_ESHAPE, "%s: expected (B,H,W,1) or (B,H,W), got rank %d", name, (int)t->ndim);
    if (t->ndim == 4 && t->shape[3] != 1)
        return fail(BTSLPG_ESHAPE, "%s: last dimension must be 1, got %lld", name, (long long)t->shape[3]);
    int64_t s[4];
    strides_of(t, s);
    v.B = t->shape[0]; v.H = t->shape[1]; v.W = t->shape[2]; v.C = 1;
    v.sB = s[0]; v.sH = s[1]; v.sW = s[2]; v.sC = 1;
    return 0;
}

inline bool is_contig_nhwc(const View &v) {
    return v.sC == 1 && v.sW == v.C && v.sH == v.W * v.C && v.sB == v.H * v.W * v.C;
}

// (B,H,W,C) NHWC view whose pixels are uniformly strided (e.g. a channel slice of a wider NHWC buffer)
inline int parse_pixel_strided(const BtsTensor *t, const char *name, View &v) {
    if (int e = parse_nhwc(t, name, v)) return e;
    const int64_t npix = v.B * v.H * v.W;
    if (npix * v.C == 0) return 0;
    if (v.C > 1 && v.sC != 1) return fail(BTSLPG_ELAYOUT, "%s: channel stride must be 1", name);
    const bool uniform = (v.H == 1 || v.sH == v.W * v.sW) && (v.B == 1 || v.sB == v.H * v.sH);
    if (!uniform) return fail(BTSLPG_ELAYOUT, "%s: pixels must be uniformly strided (a channel slice of a contiguous NHWC tensor)", name);
    if (v.sW < v.C) return fail(BTSLPG_ELAYOUT, "%s: pixel stride %lld is smaller than the channel count %lld", name, (long long)v.sW, (long long)v.C);
    return 0;
}

// flat (B,H,W[,1]) map: contiguous, 16-byte aligned
inline int parse_flat(const BtsTensor *t, const char *name, View &v, int64_t &n) {
    if (int e = parse_map(t, name, v)) return e;
    n = v.B * v.H * v.W;
    // strides of extent-1 dimensions carry no information (torch reports arbitrary values there)
    if (n > 0 && !((v.W == 1 || v.sW == 1) && (v.H == 1 || v.sH == v.W) && (v.B == 1 || v.sB == v.H * v.W)))
        return fail(BTSLPG_ELAYOUT, "%s: must be contiguous", name);
    if (!v.aligned(16)) return fail(BTSLPG_ELAYOUT, "%s: must be 16-byte aligned", name);
    return 0;
}

// float32 device vector with at least `need` elements
inline int parse_f32_vec(const BtsTensor *t, const char *name, int64_t need, int dev, float *&ptr) {
    View v;
    if (int e = parse_common(t, name, v)) return e;
    if (v.dtype != kF32) return fail(BTSLPG_EDTYPE, "%s: must be float32", name);
    if (v.dev != dev) return fail(BTSLPG_EDEVICE, "%s: on a different device", name);
    int64_t n = 1;
    for (int k = 0; k < t->ndim; ++k) n *= t->shape[k];
    if (n < need) return fail(BTSLPG_ESHAPE, "%s: needs at least %lld float32 elements, got %lld", name, (long long)need, (long long)n);
    if (t->strides && t->ndim > 0 && t->shape[t->ndim - 1] > 1 && t->strides[t->ndim - 1] != 1)
        return fail(BTSLPG_ELAYOUT, "%s: must be contiguous", name);
    if (!v.aligned(4)) return fail(BTSLPG_ELAYOUT, "%s: misaligned", name);
    ptr = reinterpret_cast<float *>(v.ptr);
    return 0;
}

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) {
            err = cudaSetDevice(dev);
            switched = (err == cudaSuccess);
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(BTSLPG_ECUDA, "%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Shape / layout checks shared by forward and backward.
// ---------------------------------------------------------------------------------------------
struct LayerGeom {
    View coef, full, ds;
    bool has_full = false, has_ds = false;
    int r = 0, d = 0;
};

inline int check_geom(LayerGeom &g, const char *coef_name, const char *full_name, const char *ds_name) {
    if (g.r < 1 || g.r > 64) return fail(BTSLPG_EINVAL, "upratio must be in [1, 64], got %d", g.r);
    if (g.coef.C != 3) return fail(BTSLPG_ESHAPE, "%s: last dimension must be 3 [phi, theta, dist], got %lld", coef_name, (long long)g.coef.C);
    const int64_t H = g.coef.H * g.r, W = g.coef.W * g.r;
    if (g.has_full) {
        if (g.full.B != g.coef.B || g.full.H != H || g.full.W != W)
            return fail(BTSLPG_ESHAPE, "%s: expected (%lld,%lld,%lld[,1]) = (B, h*%d, w*%d), got (%lld,%lld,%lld)", full_name,
                        (long long)g.coef.B, (long long)H, (long long)W, g.r, g.r, (long long)g.full.B, (long long)g.full.H, (long long)g.full.W);
        if (g.full.dtype != g.coef.dtype) return fail(BTSLPG_EDTYPE, "%s: dtype differs from %s", full_name, coef_name);
        if (g.full.dev != g.coef.dev) return fail(BTSLPG_EDEVICE, "%s: on a different device than %s", full_name, coef_name);
    }
    if (g.has_ds) {
        if (g.d < 1 || g.r % g.d != 0) return fail(BTSLPG_EINVAL, "ds_stride %d must be >= 1 and divide upratio %d", g.d, g.r);
        if (g.ds.B != g.coef.B || g.ds.H != H / g.d || g.ds.W != W / g.d)
            return fail(BTSLPG_ESHAPE, "%s: expected (%lld,%lld,%lld[,1]) = full[:, ::%d, ::%d], got (%lld,%lld,%lld)", ds_name,
                        (long long)g.coef.B, (long long)(H / g.d), (long long)(W / g.d), g.d, g.d, (long long)g.ds.B, (long long)g.ds.H, (long long)g.ds.W);
        if (g.ds.dtype != g.coef.dtype) return fail(BTSLPG_EDTYPE, "%s: dtype differs from %s", ds_name, coef_name);
        if (g.ds.dev != g.coef.dev) return fail(BTSLPG_EDEVICE, "%s: on a different device than %s", ds_name, coef_name);
    } else {
        g.d = 0;
    }
    if (g.coef.B * g.coef.H * g.coef.W >= (int64_t)1 << 31) return fail(BTSLPG_ESHAPE, "%s: more than 2^31 coarse pixels", coef_name);
    return 0;
}

inline int parse_layer_fwd(const BtsTensor *coef, int upratio, BtsTensor *out_full, BtsTensor *out_ds, int ds_stride, LayerGeom &g) {
    if (int e = parse_nhwc(coef, "coef", g.coef)) return e;
    if (int e = parse_map(out_full, "out_full", g.full)) return e;
    g.has_full = true;
    g.has_ds = out_ds != nullptr;
    if (g.has_ds) {
        if (int e = parse_map(out_ds, "out_ds", g.ds)) return e;
    }
    g.r = upratio;
    g.d = ds_stride;
    return check_geom(g, "coef", "out_full", "out_ds");
}

inline int parse_layer_bwd(const BtsTensor *coef, const BtsTensor *g_full, const BtsTensor *g_ds, int upratio, int ds_stride,
                    BtsTensor *g_coef, LayerGeom &g, View &gc) {
    if (int e = parse_nhwc(coef, "coef", g.coef)) return e;
    g.has_full = g_full != nullptr;
    g.has_ds = g_ds != nullptr;
    if (g.has_full) {
        if (int e = parse_map(g_full, "g_full", g.full)) return e;
    }
    if (g.has_ds) {
        if (int e = parse_map(g_ds, "g_ds", g.ds)) return e;
    }
    g.r = upratio;
    g.d = ds_stride;
    if (int e = check_geom(g, "coef", "g_full", "g_ds")) return e;
    if (g_coef) {
        if (int e = parse_nhwc(g_coef, "g_coef", gc)) return e;
        if (gc.B != g.coef.B || gc.H != g.coef.H || gc.W != g.coef.W || gc.C != 3)
            return fail(BTSLPG_ESHAPE, "g_coef: shape must equal coef's (B,h,w,3)");
        if (gc.dtype != g.coef.dtype) return fail(BTSLPG_EDTYPE, "g_coef: dtype differs from coef");
        if (gc.dev != g.coef.dev) return fail(BTSLPG_EDEVICE, "g_coef: on a different device than coef");
    }
    return 0;
}

// resident CTAs of a kernel on the CURRENT device (occupancy x SM count)
template <typename KernelT> inline int occupancy_blocks(KernelT kernel, int threads) {
    int per_sm = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
    if (per_sm < 1) per_sm = 1;
    if (sms < 1) sms = 148;
    return per_sm * sms;
}


template <typename KernelT> inline int occupancy_blocks_smem(KernelT kernel, int threads, int smem) {
    int per_sm = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // ask for the largest shared-memory carve-out so that as many ring-carrying CTAs as the registers allow are resident
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (per_sm < 1) per_sm = 1;
    if (sms < 1) sms = 148;
    return per_sm * sms;
}

// Function attributes and occupancy are per device: a call site caches its value per device id, so that a process that
// drives several GPUs through the ABI (DeviceGuard) opts every one of them into the large dynamic shared memory.
struct PerDevice {
    std::atomic<int> v[64] = {};
    template <typename F> int get(F &&f) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) return f();
        int x = v[dev].load(std::memory_order_relaxed);
        if (x == 0) {
            x = f();
            v[dev].store(x, std::memory_order_relaxed);
        }
        return x;
    }
};

}  // namespace btslpg_api
