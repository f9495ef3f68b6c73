// optim_api.cu -- C ABI of the fused optimizer step (include/btslpg.h: btslpg_adam_step); one translation unit of libbtslpg.so.
#include "api_common.cuh"
#include "optim_kernels.cuh"

using namespace btslpg;
using namespace btslpg_api;

namespace {

// flat float32 buffer: any rank, contiguous, 16-byte aligned
int parse_flat_f32(const BtsTensor *t, const char *name, int dev, int64_t want, float *&ptr, int64_t &n, int &dev_out) {
    View v;
    if (int e = parse_common(t, name, v)) return e;
    if (v.dtype != kF32) return fail(BTSLPG_EDTYPE, "%s: must be float32", name);
    n = 1;
    for (int k = 0; k < t->ndim; ++k) n *= t->shape[k];
    if (t->strides) {
        int64_t acc = 1;
        for (int k = t->ndim - 1; k >= 0; --k) {
            if (t->shape[k] != 1 && t->strides[k] != acc) return fail(BTSLPG_ELAYOUT, "%s: must be contiguous", name);
            acc *= t->shape[k];
        }
    }
    if (want >= 0 && n != want) return fail(BTSLPG_ESHAPE, "%s: needs %lld elements, got %lld", name, (long long)want, (long long)n);
    if (dev >= 0 && v.dev != dev) return fail(BTSLPG_EDEVICE, "%s: on a different device", name);
    if (!v.aligned(16)) return fail(BTSLPG_ELAYOUT, "%s: must be 16-byte aligned", name);
    ptr = reinterpret_cast<float *>(v.ptr);
    dev_out = v.dev;
    return 0;
}

}  // namespace

extern "C" {

int btslpg_adam_step(BtsTensor *param, BtsTensor *grad, BtsTensor *m, BtsTensor *v, BtsTensor *state, const BtsAdamConfig *cfg, int advance,
                     void *stream) {
    if (!cfg) return fail(BTSLPG_EINVAL, "cfg is NULL");
    float *pp = nullptr, *gp = nullptr, *mp = nullptr, *vp = nullptr, *sp = nullptr;
    int64_t n = 0, k = 0;
    int dev = -1, d2 = -1;
    if (int e = parse_flat_f32(param, "param", -1, -1, pp, n, dev)) return e;
    if (int e = parse_flat_f32(grad, "grad", dev, n, gp, k, d2)) return e;
    if (int e = parse_flat_f32(m, "m", dev, n, mp, k, d2)) return e;
    if (int e = parse_flat_f32(v, "v", dev, n, vp, k, d2)) return e;
    if (int e = parse_flat_f32(state, "state", dev, -1, sp, k, d2)) return e;
    if (k < 4) return fail(BTSLPG_ESHAPE, "state: needs at least 4 32-bit words, got %lld", (long long)k);
    if (!(cfg->beta1 >= 0.0f && cfg->beta1 < 1.0f && cfg->beta2 >= 0.0f && cfg->beta2 < 1.0f))
        return fail(BTSLPG_EINVAL, "beta1 / beta2 must be in [0, 1)");
    DeviceGuard guard(dev);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", dev, cudaGetErrorString(guard.err));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AdamParams p;
    p.p = pp; p.g = gp; p.m = mp; p.v = vp;
    p.state = reinterpret_cast<int *>(sp);
    p.n = (uint64_t)n;
    p.lr_start = cfg->lr_start; p.lr_end = cfg->lr_end; p.power = cfg->power;
    p.total_steps = (double)cfg->total_steps;
    p.beta1 = cfg->beta1; p.beta2 = cfg->beta2; p.epsilon = cfg->epsilon;
    p.l1 = cfg->l1; p.l2 = cfg->l2;
    p.grad_scale = cfg->grad_scale;
    p.zero_grad = cfg->zero_grad;
    if (n > 0) {
        static PerDevice per_dev;
        const int resident = per_dev.get([&] { return occupancy_blocks(adam_step_kernel, kAdamThreads); });
        int64_t blocks = (n / 4 + 2 * kAdamThreads - 1) / (2 * kAdamThreads);
        if (blocks > resident) blocks = resident;
        if (blocks < 1) blocks = 1;
        adam_step_kernel<<<(unsigned)blocks, kAdamThreads, 0, st>>>(p);
        snprintf(tl_kernel, sizeof(tl_kernel), "adam_step<f32>");
        if (int e = check_launch("btslpg_adam_step")) return e;
    }
    if (advance) {
        adam_advance_kernel<<<1, 1, 0, st>>>(p.state, p.lr_start, p.lr_end, p.power, p.total_steps);
        if (int e = check_launch("btslpg_adam_step")) return e;
    }
    return 0;
}

}  // extern "C"
