// optim_kernels.cuh -- the optimizer step of the data-parallel training loop as ONE pass over flat buffers (sm_100a).
//
// Replaces, per training step and for all decoder variables at once:
//   custom_optimizers.py:47-59   AdamW._resource_apply_dense: the decoupled decay `var -= lr * (l1*sign(var) + l2*var)`
//                                 applied BEFORE the Adam update, then tf.keras.optimizers.Adam's update
//   tf.keras Adam (TensorFlow >= 2.1, third party, not under /root/reference; algorithm as published in
//   tf.raw_ops.ResourceApplyAdam):   alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
//                                 m += (g - m) * (1 - beta1) ; v += (g*g - v) * (1 - beta2) ; var -= alpha * m / (sqrt(v) + epsilon)
//   bts_train.py:125-131          lr(step) = (start - end) * (1 - min(step, total)/total)^0.9 + end  (BatchLRScheduler,
//                                 custom_callbacks.py:46-50: evaluated at the 0-based global step before the update)
//   bts_train.py:206 (MirroredStrategy) the 1/N of the gradient average, folded in as grad_scale
// and the framework's separate zeroing of the gradient buffer (the gradient is overwritten with zeros as it is consumed).
//
// The step counter and the learning rate live in DEVICE memory (`state`), so a CUDA graph that contains this launch
// advances them on every replay without host involvement.  7 float32 per element of traffic (read p, g, m, v; write p, m, v)
// + 1 when the gradient is zeroed: an HBM-streaming pass, 16-byte accesses, two vectors in flight per thread.
#pragma once

#include "common.cuh"

namespace btslpg {

constexpr int kAdamThreads = 512;

struct AdamParams {
    float *p;
    float *g;
    float *m;
    float *v;
    int *state;               // [0] completed update steps (int32); [1] lr of the last step (float bits, informational)
    uint64_t n;
    float lr_start, lr_end, power;
    double total_steps;       // <= 0: constant learning rate lr_start
    float beta1, beta2, epsilon;
    float l1, l2;
    float grad_scale;
    int zero_grad;
};

__device__ __forceinline__ void adam_elem(float &p, float &g_io, float &m, float &v, float gs, float lr, float alpha, float omb1, float omb2,
                                          float eps, float l1, float l2, bool decay) {
    const float g = g_io * gs;
    if (decay) {
        const float sgn = (p > 0.0f) ? 1.0f : ((p < 0.0f) ? -1.0f : 0.0f);
        float d;
        if (l1 != 0.0f && l2 != 0.0f) d = l1 * sgn + l2 * p;          // custom_optimizers.py:49-54
        else if (l1 != 0.0f) d = l1 * sgn;
        else d = l2 * p;
        p = p - lr * d;                                               // :58
    }
    m = m + (g - m) * omb1;
    v = v + (g * g - v) * omb2;
    p = p - __fdiv_rn(alpha * m, sqrtf(v) + eps);
}

__global__ void __launch_bounds__(kAdamThreads) adam_step_kernel(const __grid_constant__ AdamParams prm) {
    __shared__ float s_lr, s_alpha;
    if (threadIdx.x == 0) {
        const int step = prm.state[0];                                 // 0-based global step of THIS update
        double lr = (double)prm.lr_start;
        if (prm.total_steps > 0.0) {
            const double frac = fmin((double)step, prm.total_steps) / prm.total_steps;
            lr = ((double)prm.lr_start - (double)prm.lr_end) * pow(1.0 - frac, (double)prm.power) + (double)prm.lr_end;
        }
        const float lrf = (float)lr;                                   // tf.cast(lr, tf.float32), bts_train.py:131
        const double t = (double)step + 1.0;                           // optimizer.iterations + 1
        const double alpha = (double)lrf * sqrt(1.0 - pow((double)prm.beta2, t)) / (1.0 - pow((double)prm.beta1, t));
        s_lr = lrf;
        s_alpha = (float)alpha;
    }
    __syncthreads();
    const float lr = s_lr, alpha = s_alpha;
    const float omb1 = 1.0f - prm.beta1, omb2 = 1.0f - prm.beta2;
    const bool decay = prm.l1 != 0.0f || prm.l2 != 0.0f;
    const uint64_t nvec = prm.n / 4;
    const uint64_t stride = (uint64_t)gridDim.x * kAdamThreads;
    float4 *p4 = reinterpret_cast<float4 *>(prm.p), *g4 = reinterpret_cast<float4 *>(prm.g);
    float4 *m4 = reinterpret_cast<float4 *>(prm.m), *v4 = reinterpret_cast<float4 *>(prm.v);
    for (uint64_t i = (uint64_t)blockIdx.x * kAdamThreads + threadIdx.x; i < nvec; i += 2 * stride) {
        const uint64_t i2 = i + stride;
        const bool two = i2 < nvec;
        float4 p[2], g[2], m[2], v[2];
        p[0] = p4[i]; g[0] = __ldcs(g4 + i); m[0] = m4[i]; v[0] = v4[i];
        if (two) { p[1] = p4[i2]; g[1] = __ldcs(g4 + i2); m[1] = m4[i2]; v[1] = v4[i2]; }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (q == 1 && !two) break;
            adam_elem(p[q].x, g[q].x, m[q].x, v[q].x, prm.grad_scale, lr, alpha, omb1, omb2, prm.epsilon, prm.l1, prm.l2, decay);
            adam_elem(p[q].y, g[q].y, m[q].y, v[q].y, prm.grad_scale, lr, alpha, omb1, omb2, prm.epsilon, prm.l1, prm.l2, decay);
            adam_elem(p[q].z, g[q].z, m[q].z, v[q].z, prm.grad_scale, lr, alpha, omb1, omb2, prm.epsilon, prm.l1, prm.l2, decay);
            adam_elem(p[q].w, g[q].w, m[q].w, v[q].w, prm.grad_scale, lr, alpha, omb1, omb2, prm.epsilon, prm.l1, prm.l2, decay);
            const uint64_t k = q ? i2 : i;
            p4[k] = p[q]; m4[k] = m[q]; v4[k] = v[q];
            if (prm.zero_grad) g4[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (uint64_t i = nvec * 4; i < prm.n; ++i) {
            adam_elem(prm.p[i], prm.g[i], prm.m[i], prm.v[i], prm.grad_scale, lr, alpha, omb1, omb2, prm.epsilon, prm.l1, prm.l2, decay);
            if (prm.zero_grad) prm.g[i] = 0.0f;
        }
    }
}

// after the last chunk of a step: one more completed update
__global__ void adam_advance_kernel(int *state, float lr_start, float lr_end, float power, double total_steps) {
    const int step = state[0];
    double lr = (double)lr_start;
    if (total_steps > 0.0) {
        const double frac = fmin((double)step, total_steps) / total_steps;
        lr = ((double)lr_start - (double)lr_end) * pow(1.0 - frac, (double)power) + (double)lr_end;
    }
    state[1] = __float_as_int((float)lr);
    state[0] = step + 1;
}

}  // namespace btslpg
