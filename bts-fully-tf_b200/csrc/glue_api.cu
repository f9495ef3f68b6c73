// glue_api.cu -- one translation unit of libbtslpg.so (compiled in parallel with the others by build.py).
#include "api_common.cuh"
#include "upsample_kernels.cuh"
#include "slice_kernels.cuh"

using namespace btslpg;
using namespace btslpg_api;

#include "upsample_api.inl"
#include "slice_api.inl"
