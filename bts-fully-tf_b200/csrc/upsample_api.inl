// upsample_api.inl -- C ABI for the nearest x2 up-sampling (bts_decoder.py:31, :38, :97); included by its own .cu translation unit.

namespace {

// small (B,h,w,C) and big (B,2h,2w,C) contiguous NHWC tensors of one dtype / device
int parse_upsample(const BtsTensor *small_t, const BtsTensor *big_t, const char *small_name, const char *big_name, View &sm, View &bg) {
    if (int e = parse_nhwc(small_t, small_name, sm)) return e;
    if (int e = parse_nhwc(big_t, big_name, bg)) return e;
    if (bg.B != sm.B || bg.H != 2 * sm.H || bg.W != 2 * sm.W || bg.C != sm.C)
        return fail(BTSLPG_ESHAPE, "%s: expected (%lld,%lld,%lld,%lld) = (B, 2h, 2w, C) of %s", big_name, (long long)sm.B, (long long)(2 * sm.H),
                    (long long)(2 * sm.W), (long long)sm.C, small_name);
    if (bg.dtype != sm.dtype) return fail(BTSLPG_EDTYPE, "%s: dtype differs from %s", big_name, small_name);
    if (bg.dev != sm.dev) return fail(BTSLPG_EDEVICE, "%s: on a different device than %s", big_name, small_name);
    const int64_t n = sm.B * sm.H * sm.W * sm.C;
    if (n > 0 && (!is_contig_nhwc(sm) || !is_contig_nhwc(bg))) return fail(BTSLPG_ELAYOUT, "%s / %s: must be contiguous NHWC tensors", small_name, big_name);
    if (n >= ((int64_t)1 << 31)) return fail(BTSLPG_ESHAPE, "%s: more than 2^31 elements", small_name);
    return 0;
}

template <bool BWD> int run_upsample(const View &sm, const View &bg, cudaStream_t st, const char *what) {
    const int64_t n = sm.B * sm.H * sm.W * sm.C;
    if (n == 0) return 0;
    auto go = [&](auto tag) -> int {
        using T = decltype(tag);
        const T *src = reinterpret_cast<const T *>(BWD ? bg.ptr : sm.ptr);
        T *dst = reinterpret_cast<T *>(BWD ? sm.ptr : bg.ptr);
        const bool vec = (sm.C * (int64_t)sizeof(T)) % 16 == 0 && sm.aligned(16) && bg.aligned(16);
        if (vec) {
            UpsampleParams<T> p;
            p.in = src; p.out = dst;
            p.vec_per_px = (uint32_t)(sm.C * sizeof(T) / 16);
            p.w = (uint32_t)sm.W;
            p.nvec = (uint64_t)(sm.B * sm.H * sm.W) * p.vec_per_px;
            p.div_vpp = FastDiv(p.vec_per_px);
            p.div_w = FastDiv(p.w);
            const unsigned blocks = (unsigned)((p.nvec + kUpsampleThreads - 1) / kUpsampleThreads);
            if (BWD) upsample2x_bwd_kernel<T><<<blocks, kUpsampleThreads, 0, st>>>(p);
            else upsample2x_fwd_kernel<T><<<blocks, kUpsampleThreads, 0, st>>>(p);
            snprintf(tl_kernel, sizeof(tl_kernel), "upsample2x_%s<%s,C%lld>", BWD ? "bwd" : "fwd", ElemTraits<T>::kName, (long long)sm.C);
        } else {
            UpsampleGenericParams<T> p;
            p.in = src; p.out = dst;
            p.n = (uint64_t)n; p.C = (uint32_t)sm.C; p.w = (uint32_t)sm.W;
            p.div_c = FastDiv(p.C);
            p.div_w = FastDiv(p.w);
            const unsigned blocks = (unsigned)((p.n + kUpsampleThreads - 1) / kUpsampleThreads);
            upsample2x_generic_kernel<T, BWD><<<blocks, kUpsampleThreads, 0, st>>>(p);
            snprintf(tl_kernel, sizeof(tl_kernel), "upsample2x_%s_generic<%s,C%lld>", BWD ? "bwd" : "fwd", ElemTraits<T>::kName, (long long)sm.C);
        }
        return check_launch(what);
    };
    return sm.dtype == kF32 ? go(float{}) : go(__nv_bfloat16{});
}

}  // namespace

extern "C" {

int btslpg_upsample2x_forward(const BtsTensor *in, BtsTensor *out, void *stream) {
    View sm, bg;
    if (int e = parse_upsample(in, out, "in", "out", sm, bg)) return e;
    DeviceGuard guard(sm.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", sm.dev, cudaGetErrorString(guard.err));
    return run_upsample<false>(sm, bg, static_cast<cudaStream_t>(stream), "btslpg_upsample2x_forward");
}

int btslpg_upsample2x_backward(const BtsTensor *g_out, BtsTensor *g_in, void *stream) {
    View sm, bg;
    if (int e = parse_upsample(g_in, g_out, "g_in", "g_out", sm, bg)) return e;
    DeviceGuard guard(sm.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", sm.dev, cudaGetErrorString(guard.err));
    return run_upsample<true>(sm, bg, static_cast<cudaStream_t>(stream), "btslpg_upsample2x_backward");
}

}  // extern "C"
