// depthconv_api.cu -- one translation unit of libbtslpg.so (compiled in parallel with the others by build.py).
#include "api_common.cuh"
#include "depthconv_kernels.cuh"

using namespace btslpg;
using namespace btslpg_api;

#include "depthconv_api.inl"
