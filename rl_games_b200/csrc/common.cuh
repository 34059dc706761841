// Shared device helpers for libb200rl (sm_100a).  No torch types anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200rl.h"

#define B200RL_LAUNCH_CHECK()                          \
    do {                                               \
        cudaError_t e__ = cudaGetLastError();          \
        if (e__ != cudaSuccess) return (int)e__;       \
    } while (0)

#define B200RL_EXPORT extern "C" __attribute__((visibility("default")))
static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NV doubles per thread; result valid in thread 0 (and broadcast via smem to all).
// blockDim.x must be a multiple of 32 and <= 1024.  `sm` must hold 32*NV doubles.
template <int NV>
__device__ __forceinline__ void block_sum_d(double (&v)[NV], double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sm[warp * NV + i] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = (lane < nw) ? sm[lane * NV + i] : 0.0;
            x = warp_sum(x);
            if (lane == 0) sm[i] = x;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = sm[i];
    __syncthreads();
}

// ---- Philox4x32-10 (Salmon et al.), counter-based so launches are CUDA-graph replayable ----------
struct Philox4 { uint32_t x, y, z, w; };
__device__ __forceinline__ Philox4 philox4x32_10(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t key) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}
__device__ __forceinline__ float u32_to_unit_open(uint32_t u) {   // (0,1]
    return ((float)(u >> 8) + 1.0f) * (1.0f / 16777216.0f);
}
// two standard normals from two 32-bit words (Box-Muller)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    const float u = u32_to_unit_open(a), v = u32_to_unit_open(b);
    const float r = sqrtf(-2.0f * __logf(u));
    float s, c;
    __sincosf(6.283185307179586f * v, &s, &c);
    n0 = r * c; n1 = r * s;
}

__device__ __forceinline__ float act_fwd(float x, int act) {
    switch (act) {
        case B200RL_ACT_ELU:  return x > 0.f ? x : expm1f(x);
        case B200RL_ACT_RELU: return x > 0.f ? x : 0.f;
        case B200RL_ACT_TANH: return tanhf(x);
        default: return x;
    }
}
// derivative expressed through the activation OUTPUT a
__device__ __forceinline__ float act_bwd_from_out(float a, int act) {
    switch (act) {
        case B200RL_ACT_ELU:  return a > 0.f ? 1.f : a + 1.f;
        case B200RL_ACT_RELU: return a > 0.f ? 1.f : 0.f;
        case B200RL_ACT_TANH: return 1.f - a * a;
        default: return 1.f;
    }
}

// arena chunk mapping: local row m -> element row index
__device__ __forceinline__ int64_t chunk_row(int m, int rows_per_chunk, int64_t chunk_stride) {
    const int c = m / rows_per_chunk;
    return (int64_t)c * chunk_stride + (m - c * rows_per_chunk);
}
