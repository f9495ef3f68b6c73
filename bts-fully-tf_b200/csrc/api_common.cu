// api_common.cu -- definitions of the state shared by the translation units of libbtslpg.so, and the
// introspection / tuning entry points of include/btslpg.h that only touch that state.
#include "api_common.cuh"

namespace btslpg_api {

thread_local char tl_error[512] = "";
thread_local char tl_kernel[128] = "";
std::atomic<uint64_t> g_launches{0};   // process-wide: autograd runs backward on its own thread
std::atomic<int> g_fwd_threads{0}, g_bwd_threads{0};
std::atomic<int> g_tune_head_impl{0};        // fused head forward: 0 = TMA-staged (default), 1 = register-staged loads
std::atomic<int> g_tune_concat_impl{0};      // concat forward: 0 = staged kernel (default), 1 = chunked kernel where it applies (experiment)
std::atomic<int> g_tune_depthconv_impl{0};   // last-convolution forward: 0 = tensor-core phase 1 (default), 1 = FP32-pipe phase 1

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace btslpg_api

extern "C" int btslpg_device_synchronize(int device_id) {
    using namespace btslpg_api;
    DeviceGuard guard(device_id);
    if (guard.err != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaSetDevice(%d): %s", device_id, cudaGetErrorString(guard.err));
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(BTSLPG_ECUDA, "cudaDeviceSynchronize: %s", cudaGetErrorString(e));
    return 0;
}
