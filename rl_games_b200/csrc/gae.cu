// GAE backward scan -- replaces rl_games/triton_kernels/gae_kernel.py:16-59 (_gae_kernel) and the eager
// loop :62-79.  One thread per (env, value) column; time-major loads are coalesced across envs; each
// thread issues a whole chunk of T_CHUNK independent loads (rewards, values, dones) before the
// dependent FMA chain so that enough bytes are in flight to cover HBM latency (HBM-bound op:
// 13 B per (t, env) element with u8 dones, +4 B with fused returns).
//
// Arithmetic order is the reference's, with contraction disabled (__fmul_rn/__fadd_rn), so the
// result is BIT-IDENTICAL to _pytorch_gae on the same fp32 inputs:
//   delta = (r + ((g*nv)*nnt)) - v ;  A = delta + ((gt*nnt)*A),  g = f32(gamma), gt = f32(gamma*tau)
#include "common.cuh"

namespace {

constexpr int T_CHUNK = 16;

template <typename DT> __device__ __forceinline__ float done_to_f(DT d);
template <> __device__ __forceinline__ float done_to_f<uint8_t>(uint8_t d) { return (float)d; }
template <> __device__ __forceinline__ float done_to_f<float>(float d) { return d; }

__device__ __forceinline__ float gae_step(float r, float v, float nv, float nnt, float g, float gt, float& last) {
    const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(g, nv), nnt)), v);
    last = __fadd_rn(delta, __fmul_rn(__fmul_rn(gt, nnt), last));
    return last;
}

// ---- generic strided kernel: drop-in for compute_gae (any strides / V / dones dtype) -------------
template <typename DT, typename LDT>
__global__ void __launch_bounds__(128) gae_strided_kernel(
    const float* __restrict__ rewards, const float* __restrict__ values, const DT* __restrict__ dones,
    const float* __restrict__ last_values, const LDT* __restrict__ last_dones,
    float* __restrict__ advs, float* __restrict__ returns, int H, int N, int V,
    int64_t r_st_t, int64_t r_st_e, int64_t r_st_v, int64_t v_st_t, int64_t v_st_e, int64_t v_st_v,
    int64_t d_st_t, int64_t d_st_e, int64_t a_st_t, int64_t a_st_e, int64_t a_st_v, float g, float gt) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= (int64_t)N * V) return;
    const int e = (int)(col / V), k = (int)(col - (int64_t)e * V);
    const float* rp = rewards + e * r_st_e + k * r_st_v;
    const float* vp = values + e * v_st_e + k * v_st_v;
    const DT* dp = dones + e * d_st_e;
    float* ap = advs + e * a_st_e + k * a_st_v;
    float* retp = returns ? returns + e * a_st_e + k * a_st_v : nullptr;
    float nv = last_values[(int64_t)e * V + k];
    float nnt = __fsub_rn(1.0f, done_to_f<LDT>(last_dones[e]));
    float last = 0.f;
    int t_hi = H;
    while (t_hi > 0) {
        const int t_lo = t_hi > T_CHUNK ? t_hi - T_CHUNK : 0;
        const int n = t_hi - t_lo;
        float r[T_CHUNK], v[T_CHUNK], d[T_CHUNK];
#pragma unroll
        for (int i = 0; i < T_CHUNK; ++i) {
            if (i < n) {
                const int t = t_hi - 1 - i;
                r[i] = rp[t * r_st_t]; v[i] = vp[t * v_st_t]; d[i] = done_to_f<DT>(dp[t * d_st_t]);
            }
        }
#pragma unroll
        for (int i = 0; i < T_CHUNK; ++i) {
            if (i < n) {
                const int t = t_hi - 1 - i;
                const float a = gae_step(r[i], v[i], nv, nnt, g, gt, last);
                ap[t * a_st_t] = a;
                if (retp) retp[t * a_st_t] = __fadd_rn(a, v[i]);
                nv = v[i]; nnt = __fsub_rn(1.0f, d[i]);
            }
        }
        t_hi = t_lo;
    }
}

// ---- fused fast path: contiguous [H,N], V==1, u8 dones; + returns + moment partials --------------
// partials layout per block: {n_valid, Sv, Sv2, Sr, Sr2, Sa, Sa2, 0}
template <typename DT, typename LDT, bool HAS_MASK, bool WANT_PARTIALS, bool WRITE_RET>
__global__ void __launch_bounds__(128) gae_fused_kernel(
    const float* __restrict__ rewards, const float* __restrict__ values, const DT* __restrict__ dones,
    const float* __restrict__ last_values, const LDT* __restrict__ last_dones, const float* __restrict__ mask,
    float* __restrict__ advs, float* __restrict__ returns, double* __restrict__ partials,
    int H, int N, float g, float gt) {
    __shared__ double sm[32 * 7];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    if (e < N) {
        float nv = last_values[e];
        float nnt = __fsub_rn(1.0f, done_to_f<LDT>(last_dones[e]));
        float last = 0.f;
        int t_hi = H;
        while (t_hi > 0) {
            const int t_lo = t_hi > T_CHUNK ? t_hi - T_CHUNK : 0;
            const int n = t_hi - t_lo;
            float r[T_CHUNK], v[T_CHUNK], m[T_CHUNK];
            DT d[T_CHUNK];
#pragma unroll
            for (int i = 0; i < T_CHUNK; ++i) {
                if (i < n) {
                    const int64_t idx = (int64_t)(t_hi - 1 - i) * N + e;
                    r[i] = __ldg(rewards + idx); v[i] = __ldg(values + idx); d[i] = __ldg(dones + idx);
                    if (HAS_MASK) m[i] = __ldg(mask + idx);
                }
            }
#pragma unroll
            for (int i = 0; i < T_CHUNK; ++i) {
                if (i < n) {
                    const int64_t idx = (int64_t)(t_hi - 1 - i) * N + e;
                    const float a = gae_step(r[i], v[i], nv, nnt, g, gt, last);
                    const float ret = __fadd_rn(a, v[i]);
                    advs[idx] = a;
                    if (WRITE_RET) returns[idx] = ret;
                    if (WANT_PARTIALS) {
                        const double w = HAS_MASK ? (m[i] != 0.f ? 1.0 : 0.0) : 1.0;
                        const double dv = v[i], dr = ret, da = __fsub_rn(ret, v[i]);   // advantages = returns - values (a2c_common.py:1598)
                        acc[0] += w; acc[1] += w * dv; acc[2] += w * dv * dv; acc[3] += w * dr; acc[4] += w * dr * dr;
                        acc[5] += w * da; acc[6] += w * da * da;
                    }
                    nv = v[i]; nnt = __fsub_rn(1.0f, done_to_f<DT>(d[i]));
                }
            }
            t_hi = t_lo;
        }
    }
    if (WANT_PARTIALS) {
        block_sum_d<7>(acc, sm);
        if (threadIdx.x == 0) {
            double* p = partials + (int64_t)blockIdx.x * 8;
#pragma unroll
            for (int i = 0; i < 7; ++i) p[i] = acc[i];
            p[7] = 0.0;
        }
    }
}

}  // namespace

B200RL_EXPORT int b200rl_gae_f32(const float* rewards, const float* values, const void* dones, int dones_is_u8,
                              const float* last_values, const void* last_dones, int last_dones_is_u8,
                              float* advs, float* returns, int H, int N, int V,
                              int64_t r_st_t, int64_t r_st_e, int64_t r_st_v,
                              int64_t v_st_t, int64_t v_st_e, int64_t v_st_v,
                              int64_t d_st_t, int64_t d_st_e,
                              int64_t a_st_t, int64_t a_st_e, int64_t a_st_v,
                              double gamma, double tau, void* stream) {
    if (H < 0 || N < 0 || V <= 0) return B200RL_EINVAL;
    if (H == 0 || N == 0) return B200RL_OK;
    if (!rewards || !values || !dones || !last_values || !last_dones || !advs) return B200RL_EINVAL;
    const float g = (float)gamma, gt = (float)(gamma * tau);
    const int64_t cols = (int64_t)N * V;
    const int threads = 128;
    const unsigned blocks = (unsigned)((cols + threads - 1) / threads);
    // dense time-major [H,N,1] operands (what ExperienceBuffer holds): coalesced fast path
    const bool dense = V == 1 && r_st_e == 1 && v_st_e == 1 && a_st_e == 1 && d_st_e == 1 && r_st_t == N && v_st_t == N &&
                       a_st_t == N && d_st_t == N;
    if (dense) {
#define FAST(DT, LDT)                                                                                                   \
    do {                                                                                                                \
        if (returns)                                                                                                    \
            gae_fused_kernel<DT, LDT, false, false, true><<<blocks, threads, 0, as_stream(stream)>>>(                   \
                rewards, values, (const DT*)dones, last_values, (const LDT*)last_dones, nullptr, advs, returns, nullptr, H, N, g, gt); \
        else                                                                                                            \
            gae_fused_kernel<DT, LDT, false, false, false><<<blocks, threads, 0, as_stream(stream)>>>(                  \
                rewards, values, (const DT*)dones, last_values, (const LDT*)last_dones, nullptr, advs, nullptr, nullptr, H, N, g, gt); \
    } while (0)
        if (dones_is_u8 && last_dones_is_u8) FAST(uint8_t, uint8_t);
        else if (dones_is_u8) FAST(uint8_t, float);
        else if (last_dones_is_u8) FAST(float, uint8_t);
        else FAST(float, float);
#undef FAST
        B200RL_LAUNCH_CHECK();
        return B200RL_OK;
    }
#define LAUNCH(DT, LDT)                                                                                   \
    gae_strided_kernel<DT, LDT><<<blocks, threads, 0, as_stream(stream)>>>(                               \
        rewards, values, (const DT*)dones, last_values, (const LDT*)last_dones, advs, returns, H, N, V,   \
        r_st_t, r_st_e, r_st_v, v_st_t, v_st_e, v_st_v, d_st_t, d_st_e, a_st_t, a_st_e, a_st_v, g, gt)
    if (dones_is_u8 && last_dones_is_u8) LAUNCH(uint8_t, uint8_t);
    else if (dones_is_u8) LAUNCH(uint8_t, float);
    else if (last_dones_is_u8) LAUNCH(float, uint8_t);
    else LAUNCH(float, float);
#undef LAUNCH
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}

B200RL_EXPORT int b200rl_gae_fused_f32(const float* rewards, const float* values, const uint8_t* dones,
                                    const float* last_values, const uint8_t* last_dones, const float* mask,
                                    float* advs, float* returns, double* partials, int max_partials,
                                    int H, int N, double gamma, double tau, int* n_blocks_out_host, void* stream) {
    if (H <= 0 || N <= 0) return B200RL_EINVAL;
    if (!rewards || !values || !dones || !last_values || !last_dones || !advs || !returns) return B200RL_EINVAL;
    const float g = (float)gamma, gt = (float)(gamma * tau);
    const int threads = 128;
    const int blocks = (N + threads - 1) / threads;
    if (n_blocks_out_host) *n_blocks_out_host = blocks;
    if (partials && blocks > max_partials) return B200RL_EINVAL;
    cudaStream_t s = as_stream(stream);
    if (partials) {
        if (mask) gae_fused_kernel<uint8_t, uint8_t, true, true, true><<<blocks, threads, 0, s>>>(rewards, values, dones, last_values, last_dones, mask, advs, returns, partials, H, N, g, gt);
        else gae_fused_kernel<uint8_t, uint8_t, false, true, true><<<blocks, threads, 0, s>>>(rewards, values, dones, last_values, last_dones, mask, advs, returns, partials, H, N, g, gt);
    } else {
        gae_fused_kernel<uint8_t, uint8_t, false, false, true><<<blocks, threads, 0, s>>>(rewards, values, dones, last_values, last_dones, nullptr, advs, returns, nullptr, H, N, g, gt);
    }
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}
