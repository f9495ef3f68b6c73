// slice_api.inl -- C ABI for the strided affine + activation copy (DenseASPP glue, bts_decoder.py:46-76); included by its own .cu translation unit.

extern "C" {

int btslpg_affine_act(const BtsTensor *src, const BtsTensor *scale, const BtsTensor *shift, int act, BtsTensor *dst, void *stream) {
    return btslpg_affine_act_split(src, scale, shift, act, dst, 0, 1, stream);
}

int btslpg_affine_act_split(const BtsTensor *src, const BtsTensor *scale, const BtsTensor *shift, int act, BtsTensor *dst, int split_side, int grid,
                            void *stream) {
    if (act < 0 || act > 2) return fail(BTSLPG_EINVAL, "act must be 0 (none), 1 (elu) or 2 (relu)");
    if ((scale == nullptr) != (shift == nullptr)) return fail(BTSLPG_EINVAL, "scale and shift must be given together");
    View s, d;
    if (int e = parse_pixel_strided(src, "src", s)) return e;
    if (int e = parse_pixel_strided(dst, "dst", d)) return e;
    if (split_side < 0 || split_side > 2 || grid < 1 || grid > 64) return fail(BTSLPG_EINVAL, "split_side must be 0 (none), 1 (src) or 2 (dst), 1 <= s <= 64");
    if (split_side == 0 || grid == 1) {
        split_side = 0;
        grid = 1;
    }
    // the sub-grid side is (B*s*s, H/s, W/s, C) over the other side's (B,H,W,C)
    const View &full = split_side == 1 ? d : s, &sub = split_side == 1 ? s : d;
    if (sub.B != full.B * grid * grid || sub.H * grid != full.H || sub.W * grid != full.W || sub.C != full.C)
        return fail(BTSLPG_ESHAPE, split_side ? "the sub-grid side must be (B*s*s, H/s, W/s, C) over the other side's (B,H,W,C)"
                                              : "dst: shape differs from src");
    if (d.dtype != s.dtype) return fail(BTSLPG_EDTYPE, "dst: dtype differs from src");
    if (d.dev != s.dev) return fail(BTSLPG_EDEVICE, "dst: on a different device than src");
    float *sc = nullptr, *sh = nullptr;
    if (scale) {
        if (int e = parse_f32_vec(scale, "scale", s.C, s.dev, sc)) return e;
        if (int e = parse_f32_vec(shift, "shift", s.C, s.dev, sh)) return e;
    }
    const int64_t npix = s.B * s.H * s.W;
    if (npix * s.C == 0) return 0;
    if (npix * s.C >= ((int64_t)1 << 31)) return fail(BTSLPG_ESHAPE, "src: more than 2^31 elements");
    DeviceGuard guard(s.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", s.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // a single pixel row: the pixel stride of a (1,1,1,C) tensor carries no information
    const int64_t s_src = npix == 1 ? s.C : s.sW, s_dst = npix == 1 ? d.C : d.sW;
    auto go = [&](auto tag) -> int {
        using T = decltype(tag);
        constexpr int N = 16 / (int)sizeof(T);
        const bool vec = s.C % N == 0 && s_src % N == 0 && s_dst % N == 0 && s.aligned(16) && d.aligned(16);
        SliceParams<T> p;
        p.src = reinterpret_cast<const T *>(s.ptr);
        p.dst = reinterpret_cast<T *>(d.ptr);
        p.scale = sc; p.shift = sh;
        p.C = (uint32_t)s.C;
        p.per_px = (uint32_t)(vec ? s.C / N : s.C);
        p.n = (uint64_t)npix * p.per_px;
        p.s_src = s_src; p.s_dst = s_dst;
        p.div_pp = FastDiv(p.per_px);
        p.act = act;
        p.split = split_side;
        p.s = (uint32_t)grid; p.H = (uint32_t)full.H; p.W = (uint32_t)full.W;
        p.div_hw = FastDiv((uint32_t)(full.H * full.W)); p.div_w = FastDiv((uint32_t)full.W); p.div_s = FastDiv((uint32_t)grid);
        const unsigned blocks = (unsigned)((p.n + kSliceThreads - 1) / kSliceThreads);
        if (vec) slice_affine_act_vec_kernel<T><<<blocks, kSliceThreads, 0, st>>>(p);
        else slice_affine_act_scalar_kernel<T><<<blocks, kSliceThreads, 0, st>>>(p);
        snprintf(tl_kernel, sizeof(tl_kernel), "affine_act_%s<%s,%s%s,C%u%s>", vec ? "vec" : "scalar", ElemTraits<T>::kName, sc ? "affine+" : "",
                 act == 1 ? "elu" : act == 2 ? "relu" : "id", p.C, split_side == 1 ? ",from sub-grids" : split_side == 2 ? ",to sub-grids" : "");
        return check_launch("btslpg_affine_act");
    };
    return s.dtype == kF32 ? go(float{}) : go(__nv_bfloat16{});
}

}  // extern "C"
