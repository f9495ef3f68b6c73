// head_api.inl -- C ABI for the fused reduction-head + LPG entry points (included by its own .cu translation unit).

namespace {

struct HeadGeom {
    View feat, kern, coef;
    LayerGeom lpg;     // coef view is lpg.coef
    int C = 0;
};

int parse_kernel(const BtsTensor *t, const char *name, int C, View &v, bool writable) {
    (void)writable;
    if (int e = parse_common(t, name, v)) return e;
    if (v.dtype != kF32) return fail(BTSLPG_EDTYPE, "%s: must be float32", name);
    int64_t n = 1;
    for (int k = 0; k < t->ndim; ++k) n *= t->shape[k];
    const bool ok2 = t->ndim == 2 && t->shape[0] == C && t->shape[1] == 3;
    const bool ok4 = t->ndim == 4 && t->shape[0] == 1 && t->shape[1] == 1 && t->shape[2] == C && t->shape[3] == 3;
    if (!ok2 && !ok4) return fail(BTSLPG_ESHAPE, "%s: expected [C][3] or HWIO (1,1,C,3) with C=%d", name, C);
    if (t->strides) {
        int64_t acc = 1;
        for (int k = t->ndim - 1; k >= 0; --k) {
            if (t->shape[k] != 1 && t->strides[k] != acc) return fail(BTSLPG_ELAYOUT, "%s: must be contiguous", name);
            acc *= t->shape[k];
        }
    }
    if (!v.aligned(4)) return fail(BTSLPG_ELAYOUT, "%s: misaligned", name);
    return 0;
}

bool pixels_uniform(const View &v) { return v.sH == v.W * v.sW && v.sB == v.H * v.sH; }

// alignment rules of the one-pixel-per-lane expand used by the fused kernels
bool head_maps_ok(const LayerGeom &g) {
    const int es = g.coef.esize();
    if (g.r != 2 && g.r != 4 && g.r != 8) return false;
    if (g.has_ds && (g.d != g.r / 2 || g.r == 2)) return false;
    auto fits32 = [](int64_t s) { return s >= 0 && s < ((int64_t)1 << 31); };
    if (g.has_full && !(fits32(g.full.sB) && fits32(g.full.sH))) return false;
    if (g.has_ds && !(fits32(g.ds.sB) && fits32(g.ds.sH))) return false;
    const int row_bytes = g.r * es;
    if (row_bytes % 4) return false;
    if (g.has_full) {
        const int al = row_bytes > 32 ? 32 : row_bytes;
        if (g.full.sW != 1 || !g.full.aligned(al) || (g.full.sH * es) % al || (g.full.sB * es) % al) return false;
    }
    if (g.has_ds) {
        const int al = 2 * es;
        if (al % 4) return false;
        if (g.ds.sW != 1 || !g.ds.aligned(al) || (g.ds.sH * es) % al || (g.ds.sB * es) % al) return false;
    }
    return true;
}

template <typename T, int R, int D, int M> int launch_head_fwd(const HeadFwdParams<T> &p, cudaStream_t st) {
    const int threads = 256;
    if (g_tune_head_impl.load() == 1) {
        static PerDevice per_dev; const int resident = per_dev.get([&] { return occupancy_blocks(head_lpg_fwd_kernel<T, R, D, M>, threads); });
        uint32_t blocks = (p.iters + (threads / 32) - 1) / (threads / 32);
        if (blocks > (uint32_t)resident) blocks = resident;
        head_lpg_fwd_kernel<T, R, D, M><<<blocks, threads, 0, st>>>(p);
        snprintf(tl_kernel, sizeof(tl_kernel), "head_lpg_fwd<%s,r%d,ds%d,C%d>", ElemTraits<T>::kName, R, D, 32 * M);
        return check_launch("btslpg_reduce_forward");
    }
    constexpr int smem = head_tma_smem_bytes<T, R, M, true>(256 / 32);
    static PerDevice per_dev; const int resident = per_dev.get([&] { return occupancy_blocks_smem(head_lpg_fwd_tma_kernel<T, R, D, M>, threads, smem); });
    uint32_t blocks = (p.iters + (threads / 32) - 1) / (threads / 32);
    if (blocks > (uint32_t)resident) blocks = resident;
    head_lpg_fwd_tma_kernel<T, R, D, M><<<blocks, threads, smem, st>>>(p);
    snprintf(tl_kernel, sizeof(tl_kernel), "head_lpg_fwd_tma<%s,r%d,ds%d,C%d>", ElemTraits<T>::kName, R, D, 32 * M);
    return check_launch("btslpg_reduce_forward");
}

template <typename T, int R, int D, int M> int launch_head_bwd(HeadBwdParams<T> &p, uint32_t max_blocks, cudaStream_t st) {
    const int threads = 256;
    if (g_tune_head_impl.load() == 1) {
        static PerDevice per_dev; const int resident = per_dev.get([&] { return occupancy_blocks(head_lpg_bwd_kernel<T, R, D, M>, threads); });
        uint32_t blocks = (p.iters + (threads / 32) - 1) / (threads / 32);
        if (blocks > (uint32_t)resident) blocks = resident;
        if (blocks > max_blocks) blocks = max_blocks;
        head_lpg_bwd_kernel<T, R, D, M><<<blocks, threads, 0, st>>>(p);
        snprintf(tl_kernel, sizeof(tl_kernel), "head_lpg_bwd<%s,r%d,ds%d,C%d>", ElemTraits<T>::kName, R, D, 32 * M);
        return check_launch("btslpg_reduce_backward");
    }
    if constexpr (R == 8) {
        // Two kernels for r = 8 (profiles/r02_head_bwd8.md): whole patch per lane, 32-pixel warp tiles -- the faster one per
        // pixel (0.91 of the HBM peak at B = 128) but coarse: at B = 32, 480x640 its 4800 tiles are 2.03 rounds of the ~2400
        // resident warps, i.e. three rounds paid for two; and the lane-split kernel with 8-pixel tiles (0.86 at B = 128, 8.1
        // rounds at B = 32).  The lane-split one is used while the coarse one would run fewer than 6 rounds.
        constexpr int smem8 = head_bwd8_smem_bytes<T, M>(256 / 32);
        static PerDevice per_dev8; const int resident = per_dev8.get([&] { return occupancy_blocks_smem(head_lpg_bwd8_kernel<T, D, M>, threads, smem8); });
        const int impl = g_tune_head_impl.load();        // 2 / 3 force the coarse / the lane-split kernel (A/B measurements)
        const bool fine = impl == 3 || (impl != 2 && (uint64_t)p.iters < 6ull * (uint64_t)resident * (threads / 32));
        if (fine) {
            const uint32_t tiles = (p.npix + 7) / 8;
            uint32_t blocks = (tiles + (threads / 32) - 1) / (threads / 32);
            if (blocks > (uint32_t)resident) blocks = resident;
            if (blocks > max_blocks) blocks = max_blocks;
            head_lpg_bwd8_kernel<T, D, M><<<blocks, threads, smem8, st>>>(p);
            snprintf(tl_kernel, sizeof(tl_kernel), "head_lpg_bwd8<%s,r8,ds%d,C%d>", ElemTraits<T>::kName, D, 32 * M);
            return check_launch("btslpg_reduce_backward");
        }
    }
    constexpr int smem = head_tma_smem_bytes<T, R, M, false>(256 / 32);
    static PerDevice per_dev; const int resident = per_dev.get([&] { return occupancy_blocks_smem(head_lpg_bwd_tma_kernel<T, R, D, M>, threads, smem); });
    uint32_t blocks = (p.iters + (threads / 32) - 1) / (threads / 32);
    if (blocks > (uint32_t)resident) blocks = resident;
    if (blocks > max_blocks) blocks = max_blocks;
    head_lpg_bwd_tma_kernel<T, R, D, M><<<blocks, threads, smem, st>>>(p);
    snprintf(tl_kernel, sizeof(tl_kernel), "head_lpg_bwd_tma<%s,r%d,ds%d,C%d>", ElemTraits<T>::kName, R, D, 32 * M);
    return check_launch("btslpg_reduce_backward");
}

#define BTSLPG_HEAD_DISPATCH(FN, T, r, has_ds, C, ...)                                   \
    do {                                                                                 \
        const int key_ = (r) * 1000 + (C);                                               \
        switch (key_) {                                                                  \
            case 8032: return (has_ds) ? FN<T, 8, 4, 1>(__VA_ARGS__) : FN<T, 8, 0, 1>(__VA_ARGS__); \
            case 8064: return (has_ds) ? FN<T, 8, 4, 2>(__VA_ARGS__) : FN<T, 8, 0, 2>(__VA_ARGS__); \
            case 8128: return (has_ds) ? FN<T, 8, 4, 4>(__VA_ARGS__) : FN<T, 8, 0, 4>(__VA_ARGS__); \
            case 4032: return (has_ds) ? FN<T, 4, 2, 1>(__VA_ARGS__) : FN<T, 4, 0, 1>(__VA_ARGS__); \
            case 4064: return (has_ds) ? FN<T, 4, 2, 2>(__VA_ARGS__) : FN<T, 4, 0, 2>(__VA_ARGS__); \
            case 4128: return (has_ds) ? FN<T, 4, 2, 4>(__VA_ARGS__) : FN<T, 4, 0, 4>(__VA_ARGS__); \
            case 2032: return FN<T, 2, 0, 1>(__VA_ARGS__);                               \
            case 2064: return FN<T, 2, 0, 2>(__VA_ARGS__);                               \
            case 2128: return FN<T, 2, 0, 4>(__VA_ARGS__);                               \
            default: break;                                                              \
        }                                                                                \
    } while (0)

template <typename T> int run_head_fwd_fast(const HeadGeom &hg, cudaStream_t st) {
    const LayerGeom &g = hg.lpg;
    HeadFwdParams<T> p;
    p.feat = reinterpret_cast<const T *>(hg.feat.ptr);
    p.kernel = reinterpret_cast<const float *>(hg.kern.ptr);
    p.coef_out = reinterpret_cast<T *>(g.coef.ptr);
    p.out = reinterpret_cast<T *>(g.full.ptr);
    p.ds = g.has_ds ? reinterpret_cast<T *>(g.ds.ptr) : nullptr;
    p.out_sB = (uint32_t)g.full.sB; p.out_sH = (uint32_t)g.full.sH;
    p.ds_sB = g.has_ds ? (uint32_t)g.ds.sB : 0; p.ds_sH = g.has_ds ? (uint32_t)g.ds.sH : 0;
    p.npix = (uint32_t)(g.coef.B * g.coef.H * g.coef.W);
    p.iters = (p.npix + 31) / 32;
    p.w = FastDiv((uint32_t)g.coef.W);
    p.h = FastDiv((uint32_t)g.coef.H);
    BTSLPG_HEAD_DISPATCH(launch_head_fwd, T, g.r, g.has_ds, hg.C, p, st);
    return fail(BTSLPG_ELAYOUT, "reduce_forward: no fused variant for upratio %d, C %d", g.r, hg.C);
}

template <typename T> int run_head_bwd_fast(const HeadGeom &hg, const View *gfeat, const View *gkern, const View *gcoef,
                                            void *workspace, size_t ws_bytes, cudaStream_t st) {
    const LayerGeom &g = hg.lpg;
    HeadBwdParams<T> p;
    p.feat = reinterpret_cast<const T *>(hg.feat.ptr);
    p.kernel = reinterpret_cast<const float *>(hg.kern.ptr);
    p.coef = reinterpret_cast<const T *>(g.coef.ptr);
    p.g_full = g.has_full ? reinterpret_cast<const T *>(g.full.ptr) : nullptr;
    p.g_ds = g.has_ds ? reinterpret_cast<const T *>(g.ds.ptr) : nullptr;
    p.gf_sB = g.has_full ? (uint32_t)g.full.sB : 0; p.gf_sH = g.has_full ? (uint32_t)g.full.sH : 0;
    p.gd_sB = g.has_ds ? (uint32_t)g.ds.sB : 0; p.gd_sH = g.has_ds ? (uint32_t)g.ds.sH : 0;
    p.g_feat = gfeat ? reinterpret_cast<T *>(gfeat->ptr) : nullptr;
    p.g_kernel = gkern ? reinterpret_cast<float *>(gkern->ptr) : nullptr;
    p.g_coef_out = gcoef ? reinterpret_cast<T *>(gcoef->ptr) : nullptr;
    p.counter = reinterpret_cast<unsigned int *>(workspace);
    p.partial = reinterpret_cast<float *>(static_cast<char *>(workspace) + kHeadWorkspaceHeader);
    p.npix = (uint32_t)(g.coef.B * g.coef.H * g.coef.W);
    p.iters = (p.npix + 31) / 32;
    p.w = FastDiv((uint32_t)g.coef.W);
    p.h = FastDiv((uint32_t)g.coef.H);
    uint32_t max_blocks = kHeadMaxBlocks;
    if (gkern) {
        const size_t per_block = (size_t)hg.C * 3 * sizeof(float);
        // rows: one per CTA plus one per group of kHeadGroup CTAs (head_grid_reduce); counters for <= kHeadMaxGroups groups
        if (!workspace || ws_bytes < kHeadWorkspaceHeader + 2 * per_block)
            return fail(BTSLPG_EWORKSPACE, "reduce_backward: workspace of %zu bytes is too small (need >= %zu; "
                                           "btslpg_reduce_backward_workspace_bytes gives the recommended size)",
                        ws_bytes, kHeadWorkspaceHeader + 2 * per_block);
        if ((reinterpret_cast<uintptr_t>(workspace) % 16) != 0) return fail(BTSLPG_EWORKSPACE, "reduce_backward: workspace must be 16-byte aligned");
        const size_t rows = (ws_bytes - kHeadWorkspaceHeader) / per_block;
        size_t fit = rows * kHeadGroup / (kHeadGroup + 1);
        while (fit > 1 && fit + (fit + kHeadGroup - 1) / kHeadGroup > rows) --fit;
        if (fit < 1) fit = 1;
        if (fit < max_blocks) max_blocks = (uint32_t)fit;
        if (max_blocks > (uint32_t)(kHeadMaxGroups * kHeadGroup)) max_blocks = kHeadMaxGroups * kHeadGroup;
    }
    BTSLPG_HEAD_DISPATCH(launch_head_bwd, T, g.r, g.has_ds, hg.C, p, max_blocks, st);
    return fail(BTSLPG_ELAYOUT, "reduce_backward: no fused variant for upratio %d, C %d", g.r, hg.C);
}

template <typename T> HeadGenericParams<T> make_head_generic(const HeadGeom &hg) {
    HeadGenericParams<T> p;
    memset(&p, 0, sizeof(p));
    p.feat = reinterpret_cast<const T *>(hg.feat.ptr);
    p.f_sP = hg.feat.sW; p.f_sC = hg.feat.sC;
    p.kernel = reinterpret_cast<const float *>(hg.kern.ptr);
    p.npix = hg.feat.B * hg.feat.H * hg.feat.W;
    p.C = hg.C;
    return p;
}

bool head_fast_ok(const HeadGeom &hg) {
    return (hg.C == 32 || hg.C == 64 || hg.C == 128) && is_contig_nhwc(hg.feat) && hg.feat.aligned(16) &&
           is_contig_nhwc(hg.lpg.coef) && head_maps_ok(hg.lpg);
}

}  // namespace

extern "C" {

size_t btslpg_reduce_backward_workspace_bytes(int64_t npix, int channels) {
    if (npix <= 0 || channels <= 0) return kHeadWorkspaceHeader;
    int64_t blocks = (npix + 63) / 64;     // one CTA covers at least 8 warp tiles of 8 pixels
    if (blocks > kHeadMaxGroups * kHeadGroup) blocks = kHeadMaxGroups * kHeadGroup;
    if (blocks < 1) blocks = 1;
    blocks += (blocks + kHeadGroup - 1) / kHeadGroup + 1;           // + one row per group of CTAs (two-level reduction)
    return (size_t)kHeadWorkspaceHeader + (size_t)blocks * channels * 3 * sizeof(float);
}

int btslpg_reduce_forward(const BtsTensor *feat, const BtsTensor *kernel, int upratio, BtsTensor *coef_out,
                          BtsTensor *out_full, BtsTensor *out_ds, int ds_stride, void *stream) {
    HeadGeom hg;
    if (int e = parse_nhwc(feat, "feat", hg.feat)) return e;
    hg.C = (int)hg.feat.C;
    if (hg.C < 1) return fail(BTSLPG_ESHAPE, "feat: needs at least one channel");
    if (int e = parse_kernel(kernel, "kernel", hg.C, hg.kern, false)) return e;
    if (!coef_out) return fail(BTSLPG_EINVAL, "coef_out: tensor is NULL (the sigmoid output is needed by backward)");
    if (int e = parse_layer_fwd(coef_out, upratio, out_full, out_ds, ds_stride, hg.lpg)) return e;
    const View &c = hg.lpg.coef;
    if (c.B != hg.feat.B || c.H != hg.feat.H || c.W != hg.feat.W) return fail(BTSLPG_ESHAPE, "coef_out: must be (B,h,w,3) matching feat (B,h,w,C)");
    if (c.dtype != hg.feat.dtype) return fail(BTSLPG_EDTYPE, "coef_out: dtype differs from feat");
    if (c.dev != hg.feat.dev || hg.kern.dev != hg.feat.dev) return fail(BTSLPG_EDEVICE, "feat, kernel and coef_out must be on one device");
    if (hg.feat.B * hg.feat.H * hg.feat.W == 0) return 0;
    DeviceGuard guard(hg.feat.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", hg.feat.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (head_fast_ok(hg)) return c.dtype == kF32 ? run_head_fwd_fast<float>(hg, st) : run_head_fwd_fast<__nv_bfloat16>(hg, st);

    // generic: 1x1 conv + sigmoid, then the stand-alone LPG dispatch
    if (!pixels_uniform(hg.feat)) return fail(BTSLPG_ELAYOUT, "feat: rows/batches must be uniformly strided");
    if (!is_contig_nhwc(c)) return fail(BTSLPG_ELAYOUT, "coef_out: must be contiguous");
    const int64_t npix = hg.feat.B * hg.feat.H * hg.feat.W;
    auto go = [&](auto tag) -> int {
        using T = decltype(tag);
        auto p = make_head_generic<T>(hg);
        p.coef_out = reinterpret_cast<T *>(c.ptr);
        head_fwd_generic_kernel<T><<<(unsigned)((npix + 127) / 128), 128, 0, st>>>(p);
        return check_launch("btslpg_reduce_forward");
    };
    if (int e = (c.dtype == kF32 ? go(float{}) : go(__nv_bfloat16{}))) return e;
    return btslpg_forward(coef_out, upratio, out_full, out_ds, ds_stride, stream);
}

int btslpg_reduce_backward(const BtsTensor *feat, const BtsTensor *kernel, const BtsTensor *coef, const BtsTensor *g_full,
                           const BtsTensor *g_ds, int upratio, int ds_stride, BtsTensor *g_feat, BtsTensor *g_kernel,
                           BtsTensor *g_coef_out, void *workspace, size_t workspace_bytes, void *stream) {
    HeadGeom hg;
    if (int e = parse_nhwc(feat, "feat", hg.feat)) return e;
    hg.C = (int)hg.feat.C;
    if (hg.C < 1) return fail(BTSLPG_ESHAPE, "feat: needs at least one channel");
    if (int e = parse_kernel(kernel, "kernel", hg.C, hg.kern, false)) return e;
    View gc, gfeat, gkern;
    if (int e = parse_layer_bwd(coef, g_full, g_ds, upratio, ds_stride, g_coef_out, hg.lpg, gc)) return e;
    const View &c = hg.lpg.coef;
    if (c.B != hg.feat.B || c.H != hg.feat.H || c.W != hg.feat.W) return fail(BTSLPG_ESHAPE, "coef: must be (B,h,w,3) matching feat (B,h,w,C)");
    if (c.dtype != hg.feat.dtype) return fail(BTSLPG_EDTYPE, "coef: dtype differs from feat");
    if (c.dev != hg.feat.dev || hg.kern.dev != hg.feat.dev) return fail(BTSLPG_EDEVICE, "feat, kernel and coef must be on one device");
    if (g_feat) {
        if (int e = parse_nhwc(g_feat, "g_feat", gfeat)) return e;
        if (gfeat.B != hg.feat.B || gfeat.H != hg.feat.H || gfeat.W != hg.feat.W || gfeat.C != hg.feat.C) return fail(BTSLPG_ESHAPE, "g_feat: shape must equal feat's");
        if (gfeat.dtype != hg.feat.dtype) return fail(BTSLPG_EDTYPE, "g_feat: dtype differs from feat");
        if (gfeat.dev != hg.feat.dev) return fail(BTSLPG_EDEVICE, "g_feat: on a different device than feat");
    }
    if (g_kernel) {
        if (int e = parse_kernel(g_kernel, "g_kernel", hg.C, gkern, true)) return e;
        if (gkern.dev != hg.feat.dev) return fail(BTSLPG_EDEVICE, "g_kernel: on a different device than feat");
    }
    if (!g_feat && !g_kernel && !g_coef_out) return fail(BTSLPG_EINVAL, "reduce_backward: all outputs are NULL");
    const int64_t npix = hg.feat.B * hg.feat.H * hg.feat.W;
    if (npix == 0) return 0;
    DeviceGuard guard(hg.feat.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", hg.feat.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    const bool fast = head_fast_ok(hg) && (!g_feat || (is_contig_nhwc(gfeat) && gfeat.aligned(16))) &&
                      (!g_coef_out || is_contig_nhwc(gc));
    if (fast) {
        return c.dtype == kF32
                   ? run_head_bwd_fast<float>(hg, g_feat ? &gfeat : nullptr, g_kernel ? &gkern : nullptr, g_coef_out ? &gc : nullptr, workspace, workspace_bytes, st)
                   : run_head_bwd_fast<__nv_bfloat16>(hg, g_feat ? &gfeat : nullptr, g_kernel ? &gkern : nullptr, g_coef_out ? &gc : nullptr, workspace, workspace_bytes, st);
    }

    // generic: LPG backward into g_coef_out, then the two head gradients
    if (!g_coef_out) return fail(BTSLPG_ELAYOUT, "reduce_backward: this layout (C=%d, strides) needs the generic path, which requires g_coef_out as scratch", hg.C);
    if (!pixels_uniform(hg.feat) || !is_contig_nhwc(c) || !is_contig_nhwc(gc)) return fail(BTSLPG_ELAYOUT, "reduce_backward: generic path needs uniformly strided feat and contiguous coef / g_coef_out");
    if (g_feat && !pixels_uniform(gfeat)) return fail(BTSLPG_ELAYOUT, "g_feat: rows/batches must be uniformly strided");
    if (int e = btslpg_backward(coef, g_full, g_ds, upratio, ds_stride, g_coef_out, stream)) return e;
    auto go = [&](auto tag) -> int {
        using T = decltype(tag);
        auto p = make_head_generic<T>(hg);
        p.coef = reinterpret_cast<const T *>(c.ptr);
        p.g_coef = reinterpret_cast<const T *>(gc.ptr);
        if (g_feat) {
            p.g_feat = reinterpret_cast<T *>(gfeat.ptr);
            p.gf_sP = gfeat.sW; p.gf_sC = gfeat.sC;
            head_bwd_feat_generic_kernel<T><<<(unsigned)((npix + 127) / 128), 128, 0, st>>>(p);
            if (int e = check_launch("btslpg_reduce_backward")) return e;
        }
        if (g_kernel) {
            p.g_kernel = reinterpret_cast<float *>(gkern.ptr);
            head_bwd_kernel_generic_kernel<T><<<hg.C, 256, 0, st>>>(p);
            if (int e = check_launch("btslpg_reduce_backward")) return e;
        }
        snprintf(tl_kernel, sizeof(tl_kernel), "head_bwd_generic<%s,C%d>", ElemTraits<T>::kName, hg.C);
        return 0;
    };
    return c.dtype == kF32 ? go(float{}) : go(__nv_bfloat16{});
}

}  // extern "C"
