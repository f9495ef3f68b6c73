// lpg_api.cu -- the LPG entry points and the introspection helpers of the C ABI of libbtslpg.so (declared in include/btslpg.h): argument
// validation, variant selection and kernel launches.  No allocation, no synchronisation, no
// CPU fallback: a host pointer or an unsupported layout is an error, never a slow path on the CPU.
#include "api_common.cuh"
#include "lpg_kernels.cuh"

using namespace btslpg;
using namespace btslpg_api;

namespace {


int block_threads(bool fwd, int r) {
    int t = fwd ? g_fwd_threads.load() : g_bwd_threads.load();
    (void)r;
    if (t <= 0) t = 128;
    if (t > 256) t = 256;
    t = (t + 31) / 32 * 32;
    return t;
}

// ---------------------------------------------------------------------------------------------
// Variant selection for the vectorised kernels.  A variant is (PX coarse pixels per group, ROWS
// patch rows per lane); candidates are tried in order of preference.  Requirements:
//  - r in {2,4,8}; ds absent or ds stride r/2 (the reference's two cases)
//  - coef contiguous and 16-byte aligned; maps with column stride 1; every vector access aligned
// ---------------------------------------------------------------------------------------------
struct Variant {
    int px = 0, rows = 0;
    bool ok() const { return px > 0; }
};
std::atomic<int> g_tune_r8_rows{0}, g_tune_r4_px{0}, g_tune_r2_px{0};

int candidates(int dtype, int r, bool fwd, Variant *out) {
    int n = 0;
    auto add = [&](int px, int rows) { out[n].px = px; out[n].rows = rows; ++n; };
    if (dtype == kF32) {
        if (r == 8) {
            const int t = g_tune_r8_rows.load();
            if (t == 8 || t == 2) add(1, t);
            if (fwd) { add(1, 8); add(1, 2); } else { add(1, 2); add(1, 8); }
        } else if (r == 4) {
            if (g_tune_r4_px.load() == 2) add(2, 4);
            add(1, 4); add(2, 4);
        } else if (r == 2) {
            if (g_tune_r2_px.load() == 4) add(4, 2);
            add(2, 2); add(4, 2); add(1, 2);
        }
    } else {
        if (r == 8) { add(2, 2); add(2, 4); }
        else if (r == 4) { add(2, 4); }
        else if (r == 2) { add(4, 2); add(2, 2); }
    }
    return n;
}

bool variant_fits(const LayerGeom &g, const Variant &v) {
    const int r = g.r, es = g.coef.esize(), px = v.px;
    if (g.coef.W % px) return false;
    if ((px * 3 * es) % 4) return false;                          // whole 32-bit words of coefficients
    const int row_bytes = px * r * es;                            // 32, 16 or 8
    if (row_bytes % 4) return false;
    if (g.has_full) {
        if (!g.full.aligned(row_bytes) || (g.full.sH * es) % row_bytes || (g.full.sB * es) % row_bytes) return false;
    }
    if (g.has_ds) {
        const int ds_bytes = px * 2 * es;
        if (ds_bytes % 4) return false;
        if (!g.ds.aligned(ds_bytes) || (g.ds.sH * es) % ds_bytes || (g.ds.sB * es) % ds_bytes) return false;
    }
    return true;
}

Variant pick_variant(const LayerGeom &g, bool fwd) {
    const int r = g.r;
    Variant none;
    if (r != 2 && r != 4 && r != 8) return none;
    if (g.has_ds && (g.d != r / 2 || r == 2)) return none;
    if (!is_contig_nhwc(g.coef) || !g.coef.aligned(16)) return none;
    if (g.has_full && g.full.sW != 1) return none;
    if (g.has_ds && g.ds.sW != 1) return none;
    auto fits32 = [](int64_t v) { return v >= 0 && v < ((int64_t)1 << 31); };   // kernels keep strides in 32 bits
    if (g.has_full && !(fits32(g.full.sB) && fits32(g.full.sH))) return none;
    if (g.has_ds && !(fits32(g.ds.sB) && fits32(g.ds.sH))) return none;
    Variant c[6];
    const int n = candidates(g.coef.dtype, r, fwd, c);
    for (int k = 0; k < n; ++k)
        if (variant_fits(g, c[k])) return c[k];
    return none;
}

template <typename T> LpgFwdParams<T> make_fwd_params(const LayerGeom &g, int px) {
    LpgFwdParams<T> p;
    p.coef = reinterpret_cast<const T *>(g.coef.ptr);
    p.out = reinterpret_cast<T *>(g.full.ptr);
    p.ds = g.has_ds ? reinterpret_cast<T *>(g.ds.ptr) : nullptr;
    p.out_sB = (uint32_t)g.full.sB; p.out_sH = (uint32_t)g.full.sH;
    p.ds_sB = g.has_ds ? (uint32_t)g.ds.sB : 0; p.ds_sH = g.has_ds ? (uint32_t)g.ds.sH : 0;
    const uint32_t wg = (uint32_t)(g.coef.W / px);
    p.groups = (uint32_t)(g.coef.B * g.coef.H * wg);
    p.wg = FastDiv(wg);
    p.h = FastDiv((uint32_t)g.coef.H);
    return p;
}

template <typename T> LpgBwdParams<T> make_bwd_params(const LayerGeom &g, const View &gcoef, int px) {
    LpgBwdParams<T> p;
    p.coef = reinterpret_cast<const T *>(g.coef.ptr);
    p.g_full = g.has_full ? reinterpret_cast<const T *>(g.full.ptr) : nullptr;
    p.g_ds = g.has_ds ? reinterpret_cast<const T *>(g.ds.ptr) : nullptr;
    p.g_coef = reinterpret_cast<T *>(gcoef.ptr);
    p.gf_sB = g.has_full ? (uint32_t)g.full.sB : 0; p.gf_sH = g.has_full ? (uint32_t)g.full.sH : 0;
    p.gd_sB = g.has_ds ? (uint32_t)g.ds.sB : 0; p.gd_sH = g.has_ds ? (uint32_t)g.ds.sH : 0;
    const uint32_t wg = (uint32_t)(g.coef.W / px);
    p.groups = (uint32_t)(g.coef.B * g.coef.H * wg);
    p.wg = FastDiv(wg);
    p.h = FastDiv((uint32_t)g.coef.H);
    return p;
}

template <typename T> LpgGenericParams<T> make_generic_params(const LayerGeom &g, const View *gcoef) {
    LpgGenericParams<T> p;
    memset(&p, 0, sizeof(p));
    p.coef = reinterpret_cast<const T *>(g.coef.ptr);
    p.c_sB = g.coef.sB; p.c_sH = g.coef.sH; p.c_sW = g.coef.sW; p.c_sC = g.coef.sC;
    if (g.has_full) { p.o_sB = g.full.sB; p.o_sH = g.full.sH; p.o_sW = g.full.sW; }
    if (g.has_ds) { p.d_sB = g.ds.sB; p.d_sH = g.ds.sH; p.d_sW = g.ds.sW; }
    if (gcoef) {
        p.g_full = g.has_full ? reinterpret_cast<const T *>(g.full.ptr) : nullptr;
        p.g_ds = g.has_ds ? reinterpret_cast<const T *>(g.ds.ptr) : nullptr;
        p.g_coef = reinterpret_cast<T *>(gcoef->ptr);
        p.gc_sB = gcoef->sB; p.gc_sH = gcoef->sH; p.gc_sW = gcoef->sW; p.gc_sC = gcoef->sC;
    } else {
        p.out = reinterpret_cast<T *>(g.full.ptr);
        p.ds = g.has_ds ? reinterpret_cast<T *>(g.ds.ptr) : nullptr;
    }
    p.B = g.coef.B; p.h = g.coef.H; p.w = g.coef.W;
    p.r = g.r; p.d = g.has_ds ? g.d : 0;
    return p;
}

// ---------------------------------------------------------------------------------------------
// Variant dispatch: (T, R, PX, ROWS, D) are compile-time.
// ---------------------------------------------------------------------------------------------
// A patch split over LPP = R/ROWS consecutive warps exchanges its partial sums through shared memory indexed by the
// CTA-local warp id: every group of LPP warps must lie inside ONE CTA, i.e. the CTA size is a multiple of 32*LPP.
inline int split_threads(int threads, int lpp) {
    const int unit = 32 * lpp;
    threads = (threads + unit - 1) / unit * unit;
    if (threads > 256) threads = 256 / unit * unit;
    return threads < unit ? unit : threads;
}

template <typename T, int R, int PX, int ROWS, int D>
void launch_fwd_variant(const LpgFwdParams<T> &p, int threads, cudaStream_t st) {
    threads = split_threads(threads, R / ROWS);
    const uint32_t nthreads = threads_for(p.groups, R / ROWS);
    lpg_fwd_vec_kernel<T, R, PX, ROWS, D><<<(nthreads + threads - 1) / threads, threads, 0, st>>>(p);
    snprintf(tl_kernel, sizeof(tl_kernel), "lpg_fwd_vec<%s,r%d,px%d,rows%d,ds%d>", ElemTraits<T>::kName, R, PX, ROWS, D);
}
template <typename T, int R, int PX, int ROWS, int D>
void launch_bwd_variant(const LpgBwdParams<T> &p, int threads, cudaStream_t st) {
    threads = split_threads(threads, R / ROWS);
    const uint32_t nthreads = threads_for(p.groups, R / ROWS);
    lpg_bwd_vec_kernel<T, R, PX, ROWS, D><<<(nthreads + threads - 1) / threads, threads, 0, st>>>(p);
    snprintf(tl_kernel, sizeof(tl_kernel), "lpg_bwd_vec<%s,r%d,px%d,rows%d,ds%d>", ElemTraits<T>::kName, R, PX, ROWS, D);
}

// every instantiated (R, PX, ROWS) per dtype
#define BTSLPG_VARIANTS_F32(X) X(8, 1, 2) X(8, 1, 8) X(4, 1, 4) X(4, 2, 4) X(2, 4, 2) X(2, 2, 2) X(2, 1, 2)
#define BTSLPG_VARIANTS_BF16(X) X(8, 2, 2) X(8, 2, 4) X(4, 2, 4) X(2, 4, 2) X(2, 2, 2)

template <typename T, typename P, bool FWD> void dispatch_vec(const P &p, int r, Variant v, bool has_ds, int threads, cudaStream_t st) {
#define BTSLPG_TRY(R, PX, ROWS)                                                                   \
    if (r == R && v.px == PX && v.rows == ROWS) {                                                 \
        if (has_ds) {                                                                             \
            if constexpr (R > 2) {                                                                \
                if constexpr (FWD) launch_fwd_variant<T, R, PX, ROWS, R / 2>(p, threads, st);     \
                else launch_bwd_variant<T, R, PX, ROWS, R / 2>(p, threads, st);                   \
            }                                                                                     \
        } else {                                                                                  \
            if constexpr (FWD) launch_fwd_variant<T, R, PX, ROWS, 0>(p, threads, st);             \
            else launch_bwd_variant<T, R, PX, ROWS, 0>(p, threads, st);                           \
        }                                                                                         \
        return;                                                                                   \
    }
    if constexpr (sizeof(T) == 4) { BTSLPG_VARIANTS_F32(BTSLPG_TRY) } else { BTSLPG_VARIANTS_BF16(BTSLPG_TRY) }
#undef BTSLPG_TRY
}

template <typename T> int run_forward(const LayerGeom &g, cudaStream_t st) {
    const int64_t npix = g.coef.B * g.coef.H * g.coef.W;
    if (npix == 0) return 0;
    const Variant v = pick_variant(g, true);
    if (v.ok()) {
        auto p = make_fwd_params<T>(g, v.px);
        tl_kernel[0] = 0;
        dispatch_vec<T, LpgFwdParams<T>, true>(p, g.r, v, g.has_ds, block_threads(true, g.r), st);
        if (tl_kernel[0]) return check_launch("btslpg_forward");
    }
    auto p = make_generic_params<T>(g, nullptr);
    const int threads = 128;
    lpg_fwd_generic_kernel<T><<<(unsigned)((npix + threads - 1) / threads), threads, 0, st>>>(p);
    snprintf(tl_kernel, sizeof(tl_kernel), "lpg_fwd_generic<%s,r%d,ds%d>", ElemTraits<T>::kName, g.r, g.d);
    return check_launch("btslpg_forward");
}

template <typename T> int run_backward(const LayerGeom &g, const View &gc, cudaStream_t st) {
    const int64_t npix = g.coef.B * g.coef.H * g.coef.W;
    if (npix == 0) return 0;
    Variant v = pick_variant(g, false);
    if (v.ok() && !(is_contig_nhwc(gc) && gc.aligned(16))) v = Variant();
    if (v.ok()) {
        auto p = make_bwd_params<T>(g, gc, v.px);
        tl_kernel[0] = 0;
        dispatch_vec<T, LpgBwdParams<T>, false>(p, g.r, v, g.has_ds, block_threads(false, g.r), st);
        if (tl_kernel[0]) return check_launch("btslpg_backward");
    }
    auto p = make_generic_params<T>(g, &gc);
    const int threads = 128;
    lpg_bwd_generic_kernel<T><<<(unsigned)((npix + threads - 1) / threads), threads, 0, st>>>(p);
    snprintf(tl_kernel, sizeof(tl_kernel), "lpg_bwd_generic<%s,r%d,ds%d>", ElemTraits<T>::kName, g.r, g.d);
    return check_launch("btslpg_backward");
}

// multi-launch eligibility: the default vector variant (VecCfg) with the reference's ds stride (or none)
template <typename T> bool multi_eligible_t(const LayerGeom &g, bool fwd) {
    if (g.r != 2 && g.r != 4 && g.r != 8) return false;
    Variant want;
    want.px = px_max<T>(g.r);
    want.rows = rows_default<T>(g.r, fwd);
    if (!pick_variant(g, fwd).ok()) return false;      // layout / ds-stride preconditions
    return variant_fits(g, want);
}
bool multi_eligible(const LayerGeom &g, bool fwd) {
    return g.coef.dtype == kF32 ? multi_eligible_t<float>(g, fwd) : multi_eligible_t<__nv_bfloat16>(g, fwd);
}

}  // namespace

// =============================================================================================
// exported C ABI
// =============================================================================================
extern "C" {

int btslpg_version(void) { return BTSLPG_VERSION; }
const char *btslpg_last_error(void) { return tl_error; }
const char *btslpg_last_kernel(void) { return tl_kernel; }
uint64_t btslpg_launch_count(void) { return g_launches.load(); }
void btslpg_reset_launch_count(void) { g_launches.store(0); }
void btslpg_set_block_threads(int fwd_threads, int bwd_threads) {
    g_fwd_threads.store(fwd_threads);
    g_bwd_threads.store(bwd_threads);
}
void btslpg_set_tuning(int key, int value) {
    switch (key) {
        case 0: g_fwd_threads.store(value); break;
        case 1: g_bwd_threads.store(value); break;
        case 2: g_tune_r8_rows.store(value); break;   // float32 r=8: patch rows per lane (2, 4 or 8)
        case 3: g_tune_r4_px.store(value); break;     // float32 r=4: coarse pixels per thread (1 or 2)
        case 7: g_tune_head_impl.store(value); break; // fused head forward: 0 TMA-staged, 1 register-staged
        case 10: g_tune_concat_impl.store(value); break;     // concat forward: 0 staged (default), 1 chunked where it applies
        case 9: g_tune_depthconv_impl.store(value); break;   // last-convolution forward: 0 tensor-core phase 1 (3xTF32), 1 FP32 pipe
        case 6: g_tune_r2_px.store(value); break;     // float32 r=2: coarse pixels per thread (2 or 4)
        default: break;
    }
}

const char *btslpg_status_string(int status) {
    switch (status) {
        case BTSLPG_OK: return "ok";
        case BTSLPG_EINVAL: return "invalid argument";
        case BTSLPG_EDTYPE: return "unsupported or mismatched dtype";
        case BTSLPG_ESHAPE: return "shape mismatch";
        case BTSLPG_EDEVICE: return "not a CUDA tensor / device mismatch";
        case BTSLPG_ELAYOUT: return "unsupported memory layout";
        case BTSLPG_EWORKSPACE: return "workspace missing or too small";
        case BTSLPG_ECUDA: return "CUDA error";
        default: return "unknown status";
    }
}

int btslpg_forward(const BtsTensor *coef, int upratio, BtsTensor *out_full, BtsTensor *out_ds, int ds_stride, void *stream) {
    LayerGeom g;
    if (int e = parse_layer_fwd(coef, upratio, out_full, out_ds, ds_stride, g)) return e;
    DeviceGuard guard(g.coef.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", g.coef.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return g.coef.dtype == kF32 ? run_forward<float>(g, st) : run_forward<__nv_bfloat16>(g, st);
}

int btslpg_backward(const BtsTensor *coef, const BtsTensor *g_full, const BtsTensor *g_ds, int upratio, int ds_stride,
                    BtsTensor *g_coef, void *stream) {
    if (!g_coef) return fail(BTSLPG_EINVAL, "g_coef: tensor is NULL");
    LayerGeom g;
    View gc;
    if (int e = parse_layer_bwd(coef, g_full, g_ds, upratio, ds_stride, g_coef, g, gc)) return e;
    DeviceGuard guard(g.coef.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", g.coef.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return g.coef.dtype == kF32 ? run_backward<float>(g, gc, st) : run_backward<__nv_bfloat16>(g, gc, st);
}

int btslpg_forward_multi(const BtsLpgForwardArgs *layers, int n, void *stream) {
    if (!layers || n < 1 || n > BTSLPG_MAX_MULTI) return fail(BTSLPG_EINVAL, "forward_multi: n must be in [1, %d]", BTSLPG_MAX_MULTI);
    LayerGeom g[BTSLPG_MAX_MULTI];
    bool fused = true;
    for (int k = 0; k < n; ++k) {
        if (int e = parse_layer_fwd(layers[k].coef, layers[k].upratio, layers[k].out_full, layers[k].out_ds, layers[k].ds_stride, g[k])) return e;
        fused = fused && multi_eligible(g[k], true) && g[k].coef.dtype == g[0].coef.dtype && g[k].coef.dev == g[0].coef.dev &&
                g[k].coef.B * g[k].coef.H * g[k].coef.W > 0;
        for (int j = 0; j < k; ++j) fused = fused && g[j].r != g[k].r;   // one layer per up-ratio slot
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!fused) {
        for (int k = 0; k < n; ++k)
            if (int e = btslpg_forward(layers[k].coef, layers[k].upratio, layers[k].out_full, layers[k].out_ds, layers[k].ds_stride, stream)) return e;
        return 0;
    }
    DeviceGuard guard(g[0].coef.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", g[0].coef.dev, cudaGetErrorString(guard.err));
    const int threads = kMultiThreads;
    auto go = [&](auto tag) -> int {
        using T = decltype(tag);
        LpgFwdMulti<T> m;
        memset(&m, 0, sizeof(m));
        uint32_t blocks = 0;
        for (int s = 0; s < kMultiSlots; ++s) {           // slot s <-> up-ratio 8 >> s
            for (int k = 0; k < n; ++k) {
                if (multi_slot(g[k].r) != s) continue;
                m.layer[s] = make_fwd_params<T>(g[k], px_max<T>(g[k].r));
                blocks += (threads_for(m.layer[s].groups, g[k].r / rows_default<T>(g[k].r, true)) + threads - 1) / threads;
            }
            m.block_end[s] = blocks;
        }
        lpg_fwd_multi_kernel<T><<<blocks, threads, 0, st>>>(m);
        snprintf(tl_kernel, sizeof(tl_kernel), "lpg_fwd_multi<%s,n%d>", ElemTraits<T>::kName, n);
        return check_launch("btslpg_forward_multi");
    };
    return g[0].coef.dtype == kF32 ? go(float{}) : go(__nv_bfloat16{});
}

int btslpg_backward_multi(const BtsLpgBackwardArgs *layers, int n, void *stream) {
    if (!layers || n < 1 || n > BTSLPG_MAX_MULTI) return fail(BTSLPG_EINVAL, "backward_multi: n must be in [1, %d]", BTSLPG_MAX_MULTI);
    LayerGeom g[BTSLPG_MAX_MULTI];
    View gc[BTSLPG_MAX_MULTI];
    bool fused = true;
    for (int k = 0; k < n; ++k) {
        if (!layers[k].g_coef) return fail(BTSLPG_EINVAL, "g_coef: tensor is NULL");
        if (int e = parse_layer_bwd(layers[k].coef, layers[k].g_full, layers[k].g_ds, layers[k].upratio, layers[k].ds_stride, layers[k].g_coef, g[k], gc[k])) return e;
        fused = fused && multi_eligible(g[k], false) && is_contig_nhwc(gc[k]) && gc[k].aligned(16) && g[k].coef.dtype == g[0].coef.dtype &&
                g[k].coef.dev == g[0].coef.dev && g[k].coef.B * g[k].coef.H * g[k].coef.W > 0;
        for (int j = 0; j < k; ++j) fused = fused && g[j].r != g[k].r;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!fused) {
        for (int k = 0; k < n; ++k)
            if (int e = btslpg_backward(layers[k].coef, layers[k].g_full, layers[k].g_ds, layers[k].upratio, layers[k].ds_stride, layers[k].g_coef, stream)) return e;
        return 0;
    }
    DeviceGuard guard(g[0].coef.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", g[0].coef.dev, cudaGetErrorString(guard.err));
    const int threads = kMultiThreads;
    auto go = [&](auto tag) -> int {
        using T = decltype(tag);
        LpgBwdMulti<T> m;
        memset(&m, 0, sizeof(m));
        uint32_t blocks = 0;
        for (int s = 0; s < kMultiSlots; ++s) {
            for (int k = 0; k < n; ++k) {
                if (multi_slot(g[k].r) != s) continue;
                m.layer[s] = make_bwd_params<T>(g[k], gc[k], px_max<T>(g[k].r));
                blocks += (threads_for(m.layer[s].groups, g[k].r / rows_default<T>(g[k].r, false)) + threads - 1) / threads;
            }
            m.block_end[s] = blocks;
        }
        lpg_bwd_multi_kernel<T><<<blocks, threads, 0, st>>>(m);
        snprintf(tl_kernel, sizeof(tl_kernel), "lpg_bwd_multi<%s,n%d>", ElemTraits<T>::kName, n);
        return check_launch("btslpg_backward_multi");
    };
    return g[0].coef.dtype == kF32 ? go(float{}) : go(__nv_bfloat16{});
}

}  // extern "C"
