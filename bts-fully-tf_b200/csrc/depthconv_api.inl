// depthconv_api.inl -- C ABI for the forward and the backward of the last convolution (bts_decoder.py:102); included by its own .cu translation unit.

extern "C" {

size_t btslpg_depthconv_backward_workspace_bytes(int channels) {
    if (channels < 1) channels = 1;
    return (size_t)kDcHeaderBytes + (size_t)kDcMaxBlocks * 9 * channels * sizeof(float);
}

int btslpg_depthconv_backward(const BtsTensor *x, const BtsTensor *kernel, const BtsTensor *g_out, int act_in, BtsTensor *g_x, BtsTensor *g_kernel,
                              void *workspace, size_t workspace_bytes, void *stream) {
    View xv, gv, gxv;
    if (act_in != 0 && act_in != 1) return fail(BTSLPG_EINVAL, "act_in: %d (0 none, 1 ELU)", act_in);
    if (int e = parse_nhwc(x, "x", xv)) return e;
    const int C = (int)xv.C;
    if (C != 16 && C != 32) return fail(BTSLPG_ESHAPE, "x: %d channels; the fused backward is built for C = 16 and C = 32 (F/16 of the reference's encoders)", C);
    if (!is_contig_nhwc(xv) || !xv.aligned(16)) return fail(BTSLPG_ELAYOUT, "x: must be a contiguous, 16-byte aligned NHWC tensor");
    int64_t ng = 0;
    if (int e = parse_flat(g_out, "g_out", gv, ng)) return e;
    if (gv.B != xv.B || gv.H != xv.H || gv.W != xv.W) return fail(BTSLPG_ESHAPE, "g_out: (B,H,W) differs from x");
    if (gv.dtype != xv.dtype) return fail(BTSLPG_EDTYPE, "g_out: dtype differs from x");
    if (gv.dev != xv.dev) return fail(BTSLPG_EDEVICE, "g_out: on a different device than x");
    float *w = nullptr, *gw = nullptr;
    if (!kernel) return fail(BTSLPG_EINVAL, "kernel: tensor is NULL");
    if (int e = parse_f32_vec(kernel, "kernel", 9 * C, xv.dev, w)) return e;
    if (!g_x && !g_kernel) return fail(BTSLPG_EINVAL, "depthconv_backward: both outputs are NULL");
    if (g_x) {
        if (int e = parse_nhwc(g_x, "g_x", gxv)) return e;
        if (gxv.B != xv.B || gxv.H != xv.H || gxv.W != xv.W || gxv.C != xv.C) return fail(BTSLPG_ESHAPE, "g_x: shape differs from x");
        if (gxv.dtype != xv.dtype) return fail(BTSLPG_EDTYPE, "g_x: dtype differs from x");
        if (gxv.dev != xv.dev) return fail(BTSLPG_EDEVICE, "g_x: on a different device than x");
        if (!is_contig_nhwc(gxv) || !gxv.aligned(16)) return fail(BTSLPG_ELAYOUT, "g_x: must be a contiguous, 16-byte aligned NHWC tensor");
    }
    if (g_kernel) {
        if (int e = parse_f32_vec(g_kernel, "g_kernel", 9 * C, xv.dev, gw)) return e;
        if (!workspace) return fail(BTSLPG_EWORKSPACE, "depthconv_backward: workspace is NULL");
        if (workspace_bytes < (size_t)kDcHeaderBytes + (size_t)9 * C * sizeof(float)) return fail(BTSLPG_EWORKSPACE, "depthconv_backward: workspace too small");
        if (reinterpret_cast<uintptr_t>(workspace) % 16) return fail(BTSLPG_EWORKSPACE, "depthconv_backward: workspace must be 16-byte aligned");
    }
    const int64_t npix = xv.B * xv.H * xv.W;
    if (npix == 0) return 0;
    if (npix * C >= ((int64_t)1 << 31) * 4) return fail(BTSLPG_ESHAPE, "x: too large");
    DeviceGuard guard(xv.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", xv.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto go = [&](auto tag, auto ctag, auto etag) -> int {
        using T = decltype(tag);
        constexpr int CC = decltype(ctag)::value;
        constexpr bool ELU = decltype(etag)::value != 0;
        DepthConvBwdParams<T> p;
        p.x = reinterpret_cast<const T *>(xv.ptr);
        p.g = reinterpret_cast<const T *>(gv.ptr);
        p.w = w;
        p.g_x = g_x ? reinterpret_cast<T *>(gxv.ptr) : nullptr;
        p.g_w = gw;
        p.counter = reinterpret_cast<unsigned int *>(workspace);
        p.partial = workspace ? reinterpret_cast<float *>(static_cast<char *>(workspace) + kDcHeaderBytes) : nullptr;
        p.B = (uint32_t)xv.B; p.H = (uint32_t)xv.H; p.W = (uint32_t)xv.W;
        p.col_blocks = (uint32_t)((xv.W + dc_tile_w<CC>() - 1) / dc_tile_w<CC>());
        p.items = (uint32_t)(xv.B * xv.H) * p.col_blocks;
        p.div_cb = FastDiv(p.col_blocks);
        p.div_h = FastDiv(p.H);
        p.vec_g = (sizeof(T) == 4 && xv.W % 4 == 0 && gv.aligned(16)) ? 1 : 0;
        constexpr int smem = depthconv_bwd_smem_bytes<T, CC>();
        static PerDevice per_dev; const int resident = per_dev.get([&] { return occupancy_blocks_smem(depthconv_bwd_kernel<T, CC, ELU>, kDcThreads, smem); });
        uint32_t blocks = p.items < (uint32_t)resident ? p.items : (uint32_t)resident;
        if (blocks > (uint32_t)kDcMaxBlocks) blocks = kDcMaxBlocks;
        if (gw) {
            const size_t fit = (workspace_bytes - kDcHeaderBytes) / ((size_t)9 * CC * sizeof(float));
            if (fit < blocks) blocks = (uint32_t)fit;
        }
        depthconv_bwd_kernel<T, CC, ELU><<<blocks, kDcThreads, smem, st>>>(p);
        snprintf(tl_kernel, sizeof(tl_kernel), ELU ? "depthconv_bwd<%s,C%d,elu>" : "depthconv_bwd<%s,C%d>", ElemTraits<T>::kName, CC);
        return check_launch("btslpg_depthconv_backward");
    };
    auto by_act = [&](auto tag, auto ctag) -> int { return act_in ? go(tag, ctag, IntC<1>{}) : go(tag, ctag, IntC<0>{}); };
    if (xv.dtype == kF32) return C == 32 ? by_act(float{}, IntC<32>{}) : by_act(float{}, IntC<16>{});
    return C == 32 ? by_act(__nv_bfloat16{}, IntC<32>{}) : by_act(__nv_bfloat16{}, IntC<16>{});
}

int btslpg_depthconv_forward(const BtsTensor *x, const BtsTensor *kernel, int act_in, int act_out, float out_scale, BtsTensor *y, void *stream) {
    View xv, yv;
    if (int e = parse_nhwc(x, "x", xv)) return e;
    const int C = (int)xv.C;
    if (C != 16 && C != 32) return fail(BTSLPG_ESHAPE, "x: %d channels; the fused forward is built for C = 16 and C = 32 (F/16 of the reference's encoders)", C);
    if (!is_contig_nhwc(xv) || !xv.aligned(16)) return fail(BTSLPG_ELAYOUT, "x: must be a contiguous, 16-byte aligned NHWC tensor");
    if (act_in != 0 && act_in != 1) return fail(BTSLPG_EINVAL, "act_in: %d (0 none, 1 ELU)", act_in);
    if (act_out != 0 && act_out != 1) return fail(BTSLPG_EINVAL, "act_out: %d (0 none, 1 sigmoid * out_scale)", act_out);
    int64_t ny = 0;
    if (!y) return fail(BTSLPG_EINVAL, "y: tensor is NULL");
    if (int e = parse_flat(y, "y", yv, ny)) return e;
    if (yv.B != xv.B || yv.H != xv.H || yv.W != xv.W) return fail(BTSLPG_ESHAPE, "y: (B,H,W) differs from x");
    if (yv.dtype != xv.dtype) return fail(BTSLPG_EDTYPE, "y: dtype differs from x");
    if (yv.dev != xv.dev) return fail(BTSLPG_EDEVICE, "y: on a different device than x");
    float *w = nullptr;
    if (!kernel) return fail(BTSLPG_EINVAL, "kernel: tensor is NULL");
    if (int e = parse_f32_vec(kernel, "kernel", 9 * C, xv.dev, w)) return e;
    const int64_t npix = xv.B * xv.H * xv.W;
    if (npix == 0) return 0;
    if (npix * C >= ((int64_t)1 << 31) * 4) return fail(BTSLPG_ESHAPE, "x: too large");
    if (xv.H * xv.W * C >= ((int64_t)1 << 31)) return fail(BTSLPG_ESHAPE, "x: one image has 2^31 elements or more (the kernels index inside an image with 32 bits)");
    const int64_t tiles_x = (xv.W + kDfTile - 1) / kDfTile, tiles_y = (xv.H + kDfTile - 1) / kDfTile;
    if (xv.B * tiles_x * tiles_y >= ((int64_t)1 << 31)) return fail(BTSLPG_ESHAPE, "x: too many tiles");
    DeviceGuard guard(xv.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", xv.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto go = [&](auto tag, auto ctag, auto etag) -> int {
        using T = decltype(tag);
        constexpr int CC = decltype(ctag)::value;
        constexpr bool ELU = decltype(etag)::value != 0;
        DepthConvFwdParams<T> p;
        p.x = reinterpret_cast<const T *>(xv.ptr);
        p.w = w;
        p.y = reinterpret_cast<T *>(yv.ptr);
        p.H = (uint32_t)xv.H; p.W = (uint32_t)xv.W;
        p.tiles_x = (uint32_t)tiles_x;
        p.items = (uint32_t)(xv.B * tiles_x * tiles_y);
        p.div_tx = FastDiv((uint32_t)tiles_x);
        p.div_ty = FastDiv((uint32_t)tiles_y);
        p.act_out = act_out;
        p.out_scale = out_scale;
        if (g_tune_depthconv_impl.load() == 1) {
            static PerDevice per_dev; const int resident = per_dev.get([&] { return occupancy_blocks_smem(depthconv_fwd_kernel<T, CC, ELU>, kDfThreads, 0); });
            const uint32_t blocks = p.items < (uint32_t)resident ? p.items : (uint32_t)resident;
            depthconv_fwd_kernel<T, CC, ELU><<<blocks, kDfThreads, 0, st>>>(p);
            snprintf(tl_kernel, sizeof(tl_kernel), "depthconv_fwd_fp32pipe<%s,C%d,%s>", ElemTraits<T>::kName, CC, ELU ? "elu" : "lin");
        } else {
            constexpr int smem = DcfMmaCfg<T, CC>::kSmemBytes;
            static PerDevice per_dev; const int resident = per_dev.get([&] { return occupancy_blocks_smem(depthconv_fwd_mma_kernel<T, CC, ELU>, kDfThreads, smem); });
            const uint32_t blocks = p.items < (uint32_t)resident ? p.items : (uint32_t)resident;
            depthconv_fwd_mma_kernel<T, CC, ELU><<<blocks, kDfThreads, smem, st>>>(p);
            snprintf(tl_kernel, sizeof(tl_kernel), "depthconv_fwd<%s,C%d,%s>", ElemTraits<T>::kName, CC, ELU ? "elu" : "lin");
        }
        return check_launch("btslpg_depthconv_forward");
    };
    auto by_act = [&](auto tag, auto ctag) -> int { return act_in ? go(tag, ctag, IntC<1>{}) : go(tag, ctag, IntC<0>{}); };
    if (xv.dtype == kF32) return C == 32 ? by_act(float{}, IntC<32>{}) : by_act(float{}, IntC<16>{});
    return C == 32 ? by_act(__nv_bfloat16{}, IntC<32>{}) : by_act(__nv_bfloat16{}, IntC<16>{});
}

}  // extern "C"
