// concat_api.inl -- C ABI for the fused activation + NHWC concat (bts_decoder.py:98-99, :42); included by its own .cu translation unit.

namespace {

struct ConcatGeom {
    View a, b, out;
    View plane[kConcatMaxPlanes];
    bool has_b = false;
    int np = 0;
    int64_t npix = 0;
};

int check_dense_nhwc(const View &v, const char *name, int vec) {
    if (!is_contig_nhwc(v) && v.B * v.H * v.W * v.C > 0) return fail(BTSLPG_ELAYOUT, "%s: must be a contiguous NHWC tensor", name);
    if (!v.aligned(16)) return fail(BTSLPG_ELAYOUT, "%s: must be 16-byte aligned", name);
    (void)vec;
    if (v.C < 1) return fail(BTSLPG_ESHAPE, "%s: needs at least one channel", name);
    return 0;
}

// parse (dense a, optional dense b, planes, concatenated tensor); `cat_name` names the (B,H,W,CT) tensor
int parse_concat(const BtsTensor *a, const BtsTensor *b, const BtsTensor *const *planes, int n_planes, const BtsTensor *cat, const char *a_name,
                 const char *b_name, const char *p_name, const char *cat_name, ConcatGeom &g, bool a_subpixel = false) {
    if (n_planes < 0 || n_planes > kConcatMaxPlanes) return fail(BTSLPG_EINVAL, "concat: n_planes must be in [0, %d]", kConcatMaxPlanes);
    if (n_planes > 0 && !planes) return fail(BTSLPG_EINVAL, "%s: array is NULL", p_name);
    if (int e = parse_nhwc(cat, cat_name, g.out)) return e;
    const int vec = 16 / g.out.esize();
    if (!is_contig_nhwc(g.out) && g.out.B * g.out.H * g.out.W > 0) return fail(BTSLPG_ELAYOUT, "%s: must be a contiguous NHWC tensor", cat_name);
    if (!g.out.aligned(16)) return fail(BTSLPG_ELAYOUT, "%s: must be 16-byte aligned", cat_name);
    auto same = [&](const View &v, const char *name) -> int {
        if (v.B != g.out.B || v.H != g.out.H || v.W != g.out.W) return fail(BTSLPG_ESHAPE, "%s: (B,H,W) differs from %s", name, cat_name);
        if (v.dtype != g.out.dtype) return fail(BTSLPG_EDTYPE, "%s: dtype differs from %s", name, cat_name);
        if (v.dev != g.out.dev) return fail(BTSLPG_EDEVICE, "%s: on a different device than %s", name, cat_name);
        return 0;
    };
    int64_t ct = 0;
    if (a) {
        if (int e = parse_nhwc(a, a_name, g.a)) return e;
        if (int e = check_dense_nhwc(g.a, a_name, vec)) return e;
        if (a_subpixel) {
            // (B, h, w, 4*CA) with (2h, 2w) = the concat's (H, W): fold it into the logical (B, H, W, CA) geometry
            if (g.a.B != g.out.B || 2 * g.a.H != g.out.H || 2 * g.a.W != g.out.W || g.a.C % 4)
                return fail(BTSLPG_ESHAPE, "%s: sub-pixel source must be (B, H/2, W/2, 4*CA) of %s (B,H,W,CT)", a_name, cat_name);
            if ((g.a.C / 4) % vec) return fail(BTSLPG_ESHAPE, "%s: sub-pixel source needs CA %% %d == 0", a_name, vec);
            if (g.a.dtype != g.out.dtype) return fail(BTSLPG_EDTYPE, "%s: dtype differs from %s", a_name, cat_name);
            if (g.a.dev != g.out.dev) return fail(BTSLPG_EDEVICE, "%s: on a different device than %s", a_name, cat_name);
            g.a.C /= 4; g.a.H *= 2; g.a.W *= 2;
        } else {
            if (int e = same(g.a, a_name)) return e;
        }
        ct += g.a.C;
    }
    g.has_b = b != nullptr;
    if (b) {
        if (int e = parse_nhwc(b, b_name, g.b)) return e;
        if (int e = check_dense_nhwc(g.b, b_name, vec)) return e;
        if (int e = same(g.b, b_name)) return e;
        ct += g.b.C;
    }
    g.np = n_planes;
    for (int k = 0; k < n_planes; ++k) {
        if (!planes[k]) continue;                       // backward: a NULL gradient plane is skipped
        int64_t n = 0;
        if (int e = parse_flat(planes[k], p_name, g.plane[k], n)) return e;
        if (int e = same(g.plane[k], p_name)) return e;
    }
    g.npix = g.out.B * g.out.H * g.out.W;
    return 0;
}

constexpr int kConcatMaxSmemBytes = 160 * 1024;

// pixels per tile: as many as fit the per-image budget, at least 8 (wide concats: 896 channels -> 8 pixels = 28 KB)
template <typename T> int concat_tile_px(int64_t ct, int images) {
    int64_t p = kConcatSmemBytes / (ct * (int64_t)sizeof(T));
    p = p / 8 * 8;
    if (p > 512) p = 512;
    if (p < 8 && 8 * ct * (int64_t)sizeof(T) * images <= kConcatMaxSmemBytes) p = 8;     // opt-in shared memory beyond 48 KB
    return (int)p;
}

template <typename KernelT> void concat_allow_smem(KernelT kernel, int smem) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // as many staged images per SM as the registers allow: ask for the largest shared-memory carve-out
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

template <typename KernelT> int concat_blocks(KernelT kernel, int smem, uint64_t ntiles) {
    int per_sm = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kConcatThreads, smem);
    if (per_sm < 1) per_sm = 1;
    if (sms < 1) sms = 148;
    uint64_t b = (uint64_t)per_sm * sms;
    if (b > ntiles) b = ntiles;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int btslpg_concat_forward(const BtsTensor *a, int a_subpixel, int act, const BtsTensor *scale, const BtsTensor *shift, const BtsTensor *b,
                          const BtsTensor *const *planes, int n_planes, int pad_channels, BtsTensor *out, void *stream) {
    if (!a) return fail(BTSLPG_EINVAL, "a: tensor is NULL");
    if (act != 0 && act != 1) return fail(BTSLPG_EINVAL, "act must be 0 (none) or 1 (elu)");
    if (pad_channels < 0 || pad_channels > 7) return fail(BTSLPG_EINVAL, "pad_channels must be in [0, 7]");
    if ((scale == nullptr) != (shift == nullptr)) return fail(BTSLPG_EINVAL, "scale and shift must be given together");
    ConcatGeom g;
    if (int e = parse_concat(a, b, planes, n_planes, out, "a", "b", "planes", "out", g, a_subpixel != 0)) return e;
    for (int k = 0; k < n_planes; ++k)
        if (!planes[k]) return fail(BTSLPG_EINVAL, "planes[%d]: tensor is NULL", k);
    const int64_t ct = g.a.C + (g.has_b ? g.b.C : 0) + n_planes + pad_channels;
    if (g.out.C != ct) return fail(BTSLPG_ESHAPE, "out: last dimension must be %lld (= CA + CB + n_planes + pad_channels), got %lld", (long long)ct, (long long)g.out.C);
    float *scale_ptr = nullptr, *shift_ptr = nullptr;
    if (scale) {
        if (int e = parse_f32_vec(scale, "scale", g.a.C, g.out.dev, scale_ptr)) return e;
        if (int e = parse_f32_vec(shift, "shift", g.a.C, g.out.dev, shift_ptr)) return e;
    }
    if (g.npix == 0) return 0;
    if (a_subpixel && g.npix >= ((int64_t)1 << 31)) return fail(BTSLPG_ESHAPE, "sub-pixel source: more than 2^31 pixels");
    DeviceGuard guard(g.out.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", g.out.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto go = [&](auto tag) -> int {
        using T = decltype(tag);
        const int P = concat_tile_px<T>(ct, 1);
        if (P < 8) return fail(BTSLPG_ESHAPE, "concat: %lld channels do not fit the staging buffer", (long long)ct);
        ConcatParams<T> p;
        memset(&p, 0, sizeof(p));
        p.a = reinterpret_cast<const T *>(g.a.ptr);
        p.b = g.has_b ? reinterpret_cast<const T *>(g.b.ptr) : nullptr;
        for (int k = 0; k < n_planes; ++k) p.plane[k] = reinterpret_cast<const T *>(g.plane[k].ptr);
        p.out = reinterpret_cast<T *>(g.out.ptr);
        p.npix = (uint64_t)g.npix;
        p.ca = (uint32_t)g.a.C; p.cb = g.has_b ? (uint32_t)g.b.C : 0; p.np = (uint32_t)n_planes; p.pad = (uint32_t)pad_channels; p.ct = (uint32_t)ct;
        p.scale = scale_ptr; p.shift = shift_ptr;
        p.sub_w = a_subpixel ? (uint32_t)(g.out.W / 2) : 0;
        p.div_w2 = FastDiv((uint32_t)(g.out.W ? g.out.W : 1));
        p.tile_px = (uint32_t)P;
        p.div_ca = FastDiv(p.ca);
        p.div_cb = FastDiv(p.cb ? p.cb : 1);
        p.act = act;
        p.vec = (p.ca % (16 / sizeof(T)) == 0) && (p.cb % (16 / sizeof(T)) == 0);
        // every 16-byte chunk of the output from exactly one source: the chunked kernel (no staging) CAN run; it is an
        // experiment behind tuning key 10 -- measured slower than the staged kernel on the decoder's five concats
        // (1954 us vs 1709 us per inference step at B = 32, 480x640; profiles/experiments/README.md)
        constexpr uint32_t V = 16 / sizeof(T);
        const bool chunked = g_tune_concat_impl.load() == 1 && p.ca % V == 0 && p.cb % V == 0 && (p.np + p.pad == 0 || p.np + p.pad == V) &&
                             g.a.aligned(16) && (!g.has_b || g.b.aligned(16)) && (!scale_ptr || p.ca % 4 == 0) &&
                             (!scale_ptr || (reinterpret_cast<uintptr_t>(scale_ptr) % 16 == 0 && reinterpret_cast<uintptr_t>(shift_ptr) % 16 == 0));
        if (chunked) {
            p.div_cpp = FastDiv(p.ct / V);
            const uint64_t nt = ((uint64_t)g.npix + kConcatChunkTilePx - 1) / kConcatChunkTilePx;
            if (act) concat_fwd_chunk_kernel<T, true><<<concat_blocks(concat_fwd_chunk_kernel<T, true>, 0, nt), kConcatThreads, 0, st>>>(p);
            else concat_fwd_chunk_kernel<T, false><<<concat_blocks(concat_fwd_chunk_kernel<T, false>, 0, nt), kConcatThreads, 0, st>>>(p);
            snprintf(tl_kernel, sizeof(tl_kernel), "concat_fwd_chunk<%s,%s%s%s,C%u+%u+%u+%u>", ElemTraits<T>::kName, act ? "elu" : "id", scale_ptr ? "+affine" : "",
                     p.sub_w ? "+subpixel" : "", p.ca, p.cb, p.np, p.pad);
            return check_launch("btslpg_concat_forward");
        }
        const int smem = (P * (int)ct * (int)sizeof(T) + 15) / 16 * 16 + (scale_ptr ? 2 * (int)p.ca * (int)sizeof(float) : 0);
        const uint64_t ntiles = ((uint64_t)g.npix + P - 1) / P;
        concat_allow_smem(concat_fwd_kernel<T>, smem);
        concat_fwd_kernel<T><<<concat_blocks(concat_fwd_kernel<T>, smem, ntiles), kConcatThreads, smem, st>>>(p);
        snprintf(tl_kernel, sizeof(tl_kernel), "concat_fwd<%s,%s%s%s,C%u+%u+%u+%u>", ElemTraits<T>::kName, act ? "elu" : "id", scale_ptr ? "+affine" : "",
                 p.sub_w ? "+subpixel" : "", p.ca, p.cb, p.np, p.pad);
        return check_launch("btslpg_concat_forward");
    };
    return g.out.dtype == kF32 ? go(float{}) : go(__nv_bfloat16{});
}

int btslpg_concat_backward_bn(const BtsTensor *g_out, const BtsTensor *y, int act, const BtsTensor *bn_pack, BtsTensor *g_a, int a_subpixel,
                              BtsTensor *g_b, BtsTensor *const *g_planes, int n_planes, int pad_channels, void *stream) {
    if (!g_a) return fail(BTSLPG_EINVAL, "g_a: tensor is NULL");
    const bool need_y = act != 0 || bn_pack != nullptr;
    if (act != 0 && act != 1) return fail(BTSLPG_EINVAL, "act must be 0 (none) or 1 (elu)");
    if (pad_channels < 0 || pad_channels > 7) return fail(BTSLPG_EINVAL, "pad_channels must be in [0, 7]");
    ConcatGeom g;
    if (int e = parse_concat(g_a, g_b, g_planes, n_planes, g_out, "g_a", "g_b", "g_planes", "g_out", g, a_subpixel != 0)) return e;
    const int64_t ct = g.a.C + (g.has_b ? g.b.C : 0) + n_planes + pad_channels;
    if (g.out.C != ct) return fail(BTSLPG_ESHAPE, "g_out: last dimension must be %lld (= CA + CB + n_planes + pad_channels), got %lld", (long long)ct, (long long)g.out.C);
    View yv;
    float *bn = nullptr;
    if (bn_pack) {
        if (a_subpixel) return fail(BTSLPG_EINVAL, "bn_pack: not available for the sub-pixel source");
        if (g.out.dtype != kF32) return fail(BTSLPG_EDTYPE, "bn_pack: float32 tensors only");
        if (int e = parse_f32_vec(bn_pack, "bn_pack", 8 * g.a.C, g.out.dev, bn)) return e;
    }
    if (need_y) {
        if (!y) return fail(BTSLPG_EINVAL, "y: the saved forward output is required when act != 0 or a BatchNorm pack is given");
        if (int e = parse_nhwc(y, "y", yv)) return e;
        if (yv.B != g.out.B || yv.H != g.out.H || yv.W != g.out.W || yv.C != g.out.C) return fail(BTSLPG_ESHAPE, "y: shape differs from g_out");
        if (yv.dtype != g.out.dtype) return fail(BTSLPG_EDTYPE, "y: dtype differs from g_out");
        if (yv.dev != g.out.dev) return fail(BTSLPG_EDEVICE, "y: on a different device than g_out");
        if (!is_contig_nhwc(yv) || !yv.aligned(16)) return fail(BTSLPG_ELAYOUT, "y: must be contiguous and 16-byte aligned");
    }
    if (g.npix == 0) return 0;
    if (a_subpixel && g.npix >= ((int64_t)1 << 31)) return fail(BTSLPG_ESHAPE, "sub-pixel source: more than 2^31 pixels");
    DeviceGuard guard(g.out.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", g.out.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto go = [&](auto tag) -> int {
        using T = decltype(tag);
        const int P = concat_tile_px<T>(ct, need_y ? 2 : 1);
        if (P < 8) return fail(BTSLPG_ESHAPE, "concat: %lld channels do not fit the staging buffer", (long long)ct);
        ConcatParams<T> p;
        memset(&p, 0, sizeof(p));
        p.g_out = reinterpret_cast<const T *>(g.out.ptr);
        p.y = need_y ? reinterpret_cast<const T *>(yv.ptr) : nullptr;
        p.bn = bn;
        p.g_a = reinterpret_cast<T *>(g.a.ptr);
        p.g_b = g.has_b ? reinterpret_cast<T *>(g.b.ptr) : nullptr;
        for (int k = 0; k < n_planes; ++k) p.g_plane[k] = g_planes[k] ? reinterpret_cast<T *>(g.plane[k].ptr) : nullptr;
        p.npix = (uint64_t)g.npix;
        p.ca = (uint32_t)g.a.C; p.cb = g.has_b ? (uint32_t)g.b.C : 0; p.np = (uint32_t)n_planes; p.pad = (uint32_t)pad_channels; p.ct = (uint32_t)ct;
        p.sub_w = a_subpixel ? (uint32_t)(g.out.W / 2) : 0;
        p.div_w2 = FastDiv((uint32_t)(g.out.W ? g.out.W : 1));
        p.tile_px = (uint32_t)P;
        p.div_ca = FastDiv(p.ca);
        p.div_cb = FastDiv(p.cb ? p.cb : 1);
        p.act = act;
        p.vec = (p.ca % (16 / sizeof(T)) == 0) && (p.cb % (16 / sizeof(T)) == 0);
        const int smem = (need_y ? 2 : 1) * P * (int)ct * (int)sizeof(T);
        const uint64_t ntiles = ((uint64_t)g.npix + P - 1) / P;
        concat_allow_smem(concat_bwd_kernel<T>, smem);
        concat_bwd_kernel<T><<<concat_blocks(concat_bwd_kernel<T>, smem, ntiles), kConcatThreads, smem, st>>>(p);
        snprintf(tl_kernel, sizeof(tl_kernel), "concat_bwd<%s,%s%s%s,C%u+%u+%u+%u>", ElemTraits<T>::kName, act ? "elu" : "id", bn ? "+bn" : "",
                 p.sub_w ? "+subpixel" : "", p.ca, p.cb, p.np, p.pad);
        return check_launch("btslpg_concat_backward");
    };
    return g.out.dtype == kF32 ? go(float{}) : go(__nv_bfloat16{});
}

int btslpg_concat_backward(const BtsTensor *g_out, const BtsTensor *y, int act, BtsTensor *g_a, int a_subpixel, BtsTensor *g_b,
                           BtsTensor *const *g_planes, int n_planes, int pad_channels, void *stream) {
    return btslpg_concat_backward_bn(g_out, y, act, nullptr, g_a, a_subpixel, g_b, g_planes, n_planes, pad_channels, stream);
}

}  // extern "C"
