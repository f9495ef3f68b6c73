// tail_api.inl -- C ABI for the decoder-tail entry points (silog loss, eval metrics); included by its own .cu translation unit.

namespace {

int same_as(const View &a, int64_t na, const View &ref, int64_t nref, const char *name, const char *ref_name) {
    if (na != nref || a.B != ref.B || a.H != ref.H || a.W != ref.W) return fail(BTSLPG_ESHAPE, "%s: shape differs from %s", name, ref_name);
    if (a.dtype != ref.dtype) return fail(BTSLPG_EDTYPE, "%s: dtype differs from %s", name, ref_name);
    if (a.dev != ref.dev) return fail(BTSLPG_EDEVICE, "%s: on a different device than %s", name, ref_name);
    return 0;
}

// persistent grid: one wave of resident CTAs (occupancy of this kernel x SMs), never more than the work needs
template <typename KernelT> int tail_blocks(KernelT kernel, int64_t n, int elems_per_vec) {
    const int64_t nvec = n / elems_per_vec;
    int64_t b = (nvec + 2 * kTailThreads - 1) / (2 * kTailThreads);      // two vectors per thread and iteration
    static PerDevice per_dev; const int64_t resident = per_dev.get([&] { return occupancy_blocks(kernel, kTailThreads); });   // queried once per kernel
    if (b > resident) b = resident;
    if (b > kTailMaxBlocks) b = kTailMaxBlocks;
    if (b < 1) b = 1;
    return (int)b;
}

int check_tail_ws(const void *ws, size_t bytes, const char *what) {
    if (!ws) return fail(BTSLPG_EWORKSPACE, "%s: workspace is NULL", what);
    if (bytes < kTailWorkspaceBytes) return fail(BTSLPG_EWORKSPACE, "%s: workspace needs %zu bytes, got %zu", what, (size_t)kTailWorkspaceBytes, bytes);
    if (reinterpret_cast<uintptr_t>(ws) % 16) return fail(BTSLPG_EWORKSPACE, "%s: workspace must be 16-byte aligned", what);
    return 0;
}

}  // namespace

extern "C" {

size_t btslpg_tail_workspace_bytes(void) { return kTailWorkspaceBytes; }

int btslpg_silog_forward(const BtsTensor *logit, const BtsTensor *y_true, float max_depth, float gt_threshold, BtsTensor *depth_est,
                         BtsTensor *loss, void *workspace, size_t workspace_bytes, void *stream) {
    View z, yt, y;
    int64_t n = 0, nyt = 0, nz = 0;
    if (!logit && !y_true) return fail(BTSLPG_EINVAL, "silog_forward: logit and y_true are both NULL (nothing to compute)");
    if (int e = parse_flat(depth_est, "depth_est", y, n)) return e;
    if (logit) {
        if (int e = parse_flat(logit, "logit", z, nz)) return e;
        if (int e = same_as(z, nz, y, n, "logit", "depth_est")) return e;
    }
    z.dtype = y.dtype; z.dev = y.dev;
    float *loss_ptr = nullptr;
    if (y_true) {
        if (int e = parse_flat(y_true, "y_true", yt, nyt)) return e;
        if (int e = same_as(yt, nyt, y, n, "y_true", "depth_est")) return e;
        if (!loss) return fail(BTSLPG_EINVAL, "loss: tensor is NULL (required with y_true)");
        if (int e = parse_f32_vec(loss, "loss", 1, z.dev, loss_ptr)) return e;
        if (int e = check_tail_ws(workspace, workspace_bytes, "btslpg_silog_forward")) return e;
    }
    DeviceGuard guard(z.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", z.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto go = [&](auto tag) -> int {
        using T = decltype(tag);
        SilogFwdParams<T> p;
        p.logit = logit ? reinterpret_cast<const T *>(z.ptr) : nullptr;
        p.y_true = y_true ? reinterpret_cast<const T *>(yt.ptr) : nullptr;
        p.depth = reinterpret_cast<T *>(y.ptr);
        p.loss = loss_ptr;
        p.workspace = workspace;
        p.n = (uint64_t)n;
        p.max_depth = max_depth;
        p.threshold = gt_threshold;
        silog_fwd_kernel<T><<<tail_blocks(silog_fwd_kernel<T>, n, TailVec<T>::N), kTailThreads, 0, st>>>(p);
        snprintf(tl_kernel, sizeof(tl_kernel), "silog_fwd<%s,%s>", ElemTraits<T>::kName, !y_true ? "depth" : logit ? "depth+loss" : "loss");
        return check_launch("btslpg_silog_forward");
    };
    return z.dtype == kF32 ? go(float{}) : go(__nv_bfloat16{});
}

int btslpg_silog_backward(const BtsTensor *depth_est, const BtsTensor *y_true, float max_depth, float gt_threshold, const BtsTensor *g_loss,
                          const void *workspace, size_t workspace_bytes, int wrt_logit, BtsTensor *g_out, void *stream) {
    View y, yt, g;
    int64_t n = 0, nyt = 0, ng = 0;
    if (int e = parse_flat(depth_est, "depth_est", y, n)) return e;
    if (int e = parse_flat(y_true, "y_true", yt, nyt)) return e;
    if (int e = same_as(yt, nyt, y, n, "y_true", "depth_est")) return e;
    if (int e = parse_flat(g_out, "g_out", g, ng)) return e;
    if (int e = same_as(g, ng, y, n, "g_out", "depth_est")) return e;
    float *gl = nullptr;
    if (g_loss)
        if (int e = parse_f32_vec(g_loss, "g_loss", 1, y.dev, gl)) return e;
    if (int e = check_tail_ws(workspace, workspace_bytes, "btslpg_silog_backward")) return e;
    DeviceGuard guard(y.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", y.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto go = [&](auto tag) -> int {
        using T = decltype(tag);
        SilogBwdParams<T> p;
        p.depth = reinterpret_cast<const T *>(y.ptr);
        p.y_true = reinterpret_cast<const T *>(yt.ptr);
        p.g_loss = gl;
        p.workspace = workspace;
        p.g_logit = reinterpret_cast<T *>(g.ptr);
        p.n = (uint64_t)n;
        p.max_depth = max_depth;
        p.threshold = gt_threshold;
        p.wrt_logit = wrt_logit ? 1 : 0;
        silog_bwd_kernel<T><<<tail_blocks(silog_bwd_kernel<T>, n, TailVec<T>::N), kTailThreads, 0, st>>>(p);
        snprintf(tl_kernel, sizeof(tl_kernel), "silog_bwd<%s,%s>", ElemTraits<T>::kName, wrt_logit ? "logit" : "depth");
        return check_launch("btslpg_silog_backward");
    };
    return y.dtype == kF32 ? go(float{}) : go(__nv_bfloat16{});
}

int btslpg_eval_metrics_png16(const BtsTensor *y_true, const BtsTensor *y_pred, float min_depth_eval, float max_depth_eval, BtsTensor *metrics,
                              float png_max_depth, BtsTensor *png, void *workspace, size_t workspace_bytes, void *stream) {
    View yt, yp;
    int64_t n = 0, np_ = 0;
    const bool with_metrics = y_true != nullptr;
    if (int e = parse_flat(y_pred, "y_pred", yp, np_)) return e;
    if (with_metrics) {
        if (int e = parse_flat(y_true, "y_true", yt, n)) return e;
        if (int e = same_as(yp, np_, yt, n, "y_pred", "y_true")) return e;
    } else if (!png) {
        return fail(BTSLPG_EINVAL, "y_true and png are both NULL: nothing to compute");
    }
    float *out = nullptr;
    if (with_metrics) {
        if (!metrics) return fail(BTSLPG_EINVAL, "metrics: tensor is NULL");
        if (int e = parse_f32_vec(metrics, "metrics", 10, yp.dev, out)) return e;
        if (int e = check_tail_ws(workspace, workspace_bytes, "btslpg_eval_metrics")) return e;
    }
    uint16_t *png_ptr = nullptr;
    if (png) {
        // uint16 image (kDLUInt = 1, 16 bits), same number of elements as y_pred, contiguous, 16-byte aligned
        if (png->device.device_type != 2 && png->device.device_type != 13)
            return fail(BTSLPG_EDEVICE, "png: not a CUDA tensor; host tensors are not accepted -- there is no CPU fallback");
        if (png->dtype.code != 1 || png->dtype.bits != 16 || png->dtype.lanes != 1) return fail(BTSLPG_EDTYPE, "png: must be uint16");
        if (png->device.device_id != yp.dev) return fail(BTSLPG_EDEVICE, "png: on a different device than y_pred");
        int64_t m = 1, acc = 1;
        for (int k = 0; k < png->ndim; ++k) m *= png->shape[k];
        if (m != np_) return fail(BTSLPG_ESHAPE, "png: needs %lld elements like y_pred, got %lld", (long long)np_, (long long)m);
        if (png->strides) {
            for (int k = png->ndim - 1; k >= 0; --k) {
                if (png->shape[k] != 1 && png->strides[k] != acc) return fail(BTSLPG_ELAYOUT, "png: must be contiguous");
                acc *= png->shape[k];
            }
        }
        png_ptr = reinterpret_cast<uint16_t *>(static_cast<char *>(png->data) + png->byte_offset);
        if (reinterpret_cast<uintptr_t>(png_ptr) % 16) return fail(BTSLPG_ELAYOUT, "png: must be 16-byte aligned");
        if (!(png_max_depth > 0.0f)) return fail(BTSLPG_EINVAL, "png_max_depth must be positive");
    }
    n = np_;
    DeviceGuard guard(yp.dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", yp.dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto go = [&](auto tag) -> int {
        using T = decltype(tag);
        MetricsParams<T> p;
        p.y_true = with_metrics ? reinterpret_cast<const T *>(yt.ptr) : nullptr;
        p.y_pred = reinterpret_cast<const T *>(yp.ptr);
        p.out = out;
        p.workspace = workspace;
        p.n = (uint64_t)n;
        p.min_depth = min_depth_eval;
        p.max_depth = max_depth_eval;
        p.png = png_ptr;
        p.png_max_depth = png_max_depth;
        if (with_metrics && png_ptr) {
            eval_metrics_kernel<T, true, true><<<tail_blocks(eval_metrics_kernel<T, true, true>, n, TailVec<T>::N), kTailThreads, 0, st>>>(p);
            snprintf(tl_kernel, sizeof(tl_kernel), "eval_metrics_png16<%s>", ElemTraits<T>::kName);
        } else if (with_metrics) {
            eval_metrics_kernel<T, false, true><<<tail_blocks(eval_metrics_kernel<T, false, true>, n, TailVec<T>::N), kTailThreads, 0, st>>>(p);
            snprintf(tl_kernel, sizeof(tl_kernel), "eval_metrics<%s>", ElemTraits<T>::kName);
        } else {
            eval_metrics_kernel<T, true, false><<<tail_blocks(eval_metrics_kernel<T, true, false>, n, TailVec<T>::N), kTailThreads, 0, st>>>(p);
            snprintf(tl_kernel, sizeof(tl_kernel), "depth_png16<%s>", ElemTraits<T>::kName);
        }
        return check_launch("btslpg_eval_metrics");
    };
    return yp.dtype == kF32 ? go(float{}) : go(__nv_bfloat16{});
}

int btslpg_eval_metrics(const BtsTensor *y_true, const BtsTensor *y_pred, float min_depth_eval, float max_depth_eval, BtsTensor *metrics,
                        void *workspace, size_t workspace_bytes, void *stream) {
    if (!y_true) return fail(BTSLPG_EINVAL, "y_true: tensor is NULL");
    return btslpg_eval_metrics_png16(y_true, y_pred, min_depth_eval, max_depth_eval, metrics, 1.0f, nullptr, workspace, workspace_bytes, stream);
}

}  // extern "C"
