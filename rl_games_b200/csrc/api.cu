// Library identity + tiny utility kernels.
#include "common.cuh"


int g_b200rl_pdl = 0;

B200RL_EXPORT int b200rl_version(void) { return 100; }
B200RL_EXPORT int b200rl_built_arch(void) { return 100; }
B200RL_EXPORT int b200rl_set_pdl(int enable) {
    const int prev = g_b200rl_pdl;
    g_b200rl_pdl = enable ? 1 : 0;
    return prev;
}

namespace {
__global__ void fill_u32_kernel(uint32_t* p, int64_t n, uint32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
}  // namespace

B200RL_EXPORT int b200rl_fill_u32(uint32_t* p, int64_t n, uint32_t v, void* stream) {
    if (!p || n < 0) return B200RL_EINVAL;
    if (n == 0) return B200RL_OK;
    fill_u32_kernel<<<148 * 8, 256, 0, as_stream(stream)>>>(p, n, v);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}
