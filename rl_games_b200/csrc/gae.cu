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

// A/B switch between the TMA-staged kernel (default) and the register-chunk kernel (gae_fused_kernel); process-wide
static int g_gae_tma = 1;
B200RL_EXPORT int b200rl_gae_set_tma(int enable) { const int prev = g_gae_tma; g_gae_tma = enable ? 1 : 0; return prev; }
static inline int b200rl_gae_use_tma() { return g_gae_tma; }

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

// ---- TMA-staged fused kernel (the default for dense [H,N] arenas with N % 16 == 0) ----------------------------------------------
// One CTA owns a tile of E envs x all H steps.  The rows of the tile (E*4 contiguous bytes per step for rewards / values / mask,
// E bytes for dones) are pulled into shared memory by 1-D bulk copies (cp.async.bulk, mbarrier transaction counts) issued by the
// lanes of warp 0 in the first cycles of the CTA -- every byte of the tile is requested before any thread touches data, the
// dependent scan then runs out of shared memory (bank-conflict free: thread = env), advantages / returns overwrite the staged
// rewards / values in place and leave as bulk stores.  TC-step chunks through a two-stage ring cover any horizon; H <= 2*TC has
// the whole tile in flight at once.  Same arithmetic as gae_fused_kernel (gae_step): bit-identical to _pytorch_gae.
__device__ __forceinline__ uint32_t gae_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gae_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gae_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gae_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gae_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gae_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "GAE_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GAE_DONE;\n\t"
        "bra GAE_WAIT;\n\t"
        "GAE_DONE:\n\t}" ::"r"(gae_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void gae_bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(gae_smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(gae_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void gae_bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(gae_smem_u32(ssrc)), "r"(bytes) : "memory");
}

constexpr int GAE_TC = 16;      // steps per ring stage

template <int E, bool HAS_MASK, bool WANT_PARTIALS, bool WRITE_RET>
__global__ void __launch_bounds__(E) gae_tma_kernel(
    const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
    const float* __restrict__ last_values, const uint8_t* __restrict__ last_dones, const float* __restrict__ mask,
    float* __restrict__ advs, float* __restrict__ returns, double* __restrict__ partials, int H, int N, float g, float gt) {
    constexpr int TC = GAE_TC;
    constexpr int STAGE = TC * E * (8 + (HAS_MASK ? 4 : 0) + 1);
    extern __shared__ __align__(128) uint8_t gsm[];
    __shared__ uint64_t bars[2];
    __shared__ double sm[32 * 7];
    const int tid = threadIdx.x;
    const int e0 = blockIdx.x * E;
    const int ne = min(E, N - e0);                       // multiple of 16 (host-checked): every bulk copy is 16-byte granular
    const int nchunks = (H + TC - 1) / TC;
    auto stage_r = [&](int s) { return reinterpret_cast<float*>(gsm + s * STAGE); };
    auto stage_v = [&](int s) { return reinterpret_cast<float*>(gsm + s * STAGE + TC * E * 4); };
    auto stage_m = [&](int s) { return reinterpret_cast<float*>(gsm + s * STAGE + TC * E * 8); };
    auto stage_d = [&](int s) { return gsm + s * STAGE + TC * E * (8 + (HAS_MASK ? 4 : 0)); };
    // chunk c = steps [t_lo, t_hi) counted from the END of the horizon (the scan runs backwards)
    auto issue = [&](int c) {            // called by all lanes of warp 0
        const int s = c & 1;
        const int t_hi = H - c * TC, t_lo = max(t_hi - TC, 0), rows = t_hi - t_lo;
        if (tid == 0) gae_mbar_expect_tx(&bars[s], (uint32_t)rows * (uint32_t)ne * (8u + (HAS_MASK ? 4u : 0u) + 1u));
        __syncwarp();
        for (int i = tid; i < rows; i += 32) {
            const int64_t off = (int64_t)(t_lo + i) * N + e0;
            gae_bulk_g2s(stage_r(s) + i * E, rewards + off, (uint32_t)ne * 4u, &bars[s]);
            gae_bulk_g2s(stage_v(s) + i * E, values + off, (uint32_t)ne * 4u, &bars[s]);
            if (HAS_MASK) gae_bulk_g2s(stage_m(s) + i * E, mask + off, (uint32_t)ne * 4u, &bars[s]);
            gae_bulk_g2s(stage_d(s) + i * E, dones + off, (uint32_t)ne, &bars[s]);
        }
    };
    if (tid == 0) {
        gae_mbar_init(&bars[0], 1); gae_mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32) {
        issue(0);
        if (nchunks > 1) issue(1);
    }
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    const bool live = tid < ne;
    float nv = 0.f, nnt = 0.f, last = 0.f;
    if (live) {
        nv = __ldg(last_values + e0 + tid);
        nnt = __fsub_rn(1.0f, (float)__ldg(last_dones + e0 + tid));
    }
    for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1;
        const int t_hi = H - c * TC, t_lo = max(t_hi - TC, 0), rows = t_hi - t_lo;
        gae_mbar_wait(&bars[s], (uint32_t)(c >> 1) & 1u);
        float* sr = stage_r(s); float* sv = stage_v(s); const float* smk = stage_m(s); const uint8_t* sd = stage_d(s);
        if (live) {
#pragma unroll 4
            for (int i = rows - 1; i >= 0; --i) {
                const float r = sr[i * E + tid], v = sv[i * E + tid];
                const float a = gae_step(r, v, nv, nnt, g, gt, last);
                const float ret = __fadd_rn(a, v);
                sr[i * E + tid] = a;
                if (WRITE_RET) sv[i * E + tid] = ret;
                if (WANT_PARTIALS) {
                    const double w = HAS_MASK ? (smk[i * E + tid] != 0.f ? 1.0 : 0.0) : 1.0;
                    const double dv = v, dr = ret, da = __fsub_rn(ret, v);
                    acc[0] += w; acc[1] += w * dv; acc[2] += w * dv * dv; acc[3] += w * dr; acc[4] += w * dr * dr;
                    acc[5] += w * da; acc[6] += w * da * da;
                }
                nv = v; nnt = __fsub_rn(1.0f, (float)sd[i * E + tid]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the bulk stores
        __syncthreads();
        if (tid < 32) {
            for (int i = tid; i < rows; i += 32) {
                const int64_t off = (int64_t)(t_lo + i) * N + e0;
                gae_bulk_s2g(advs + off, sr + i * E, (uint32_t)ne * 4u);
                if (WRITE_RET) gae_bulk_s2g(returns + off, sv + i * E, (uint32_t)ne * 4u);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (c + 2 < nchunks) {
                // the stage is refilled for chunk c + 2: its stores must have finished READING shared memory first
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                issue(c + 2);
            }
        }
    }
    if (tid < 32) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // smem must outlive the last stores' reads
    if (WANT_PARTIALS) {
        block_sum_d<7>(acc, sm);
        if (tid == 0) {
            double* p = partials + (int64_t)blockIdx.x * 8;
#pragma unroll
            for (int i = 0; i < 7; ++i) p[i] = acc[i];
            p[7] = 0.0;
        }
    }
}

template <int E, bool HAS_MASK, bool WANT_PARTIALS, bool WRITE_RET>
static cudaError_t launch_gae_tma(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                                  const uint8_t* last_dones, const float* mask, float* advs, float* returns, double* partials, int H, int N,
                                  float g, float gt, cudaStream_t s) {
    constexpr int STAGE = GAE_TC * E * (8 + (HAS_MASK ? 4 : 0) + 1);
    const int blocks = (N + E - 1) / E;
    if (2 * STAGE > 48 * 1024) {
        static bool raised = false;      // per instantiation
        if (!raised) {
            cudaError_t e = cudaFuncSetAttribute(gae_tma_kernel<E, HAS_MASK, WANT_PARTIALS, WRITE_RET>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * STAGE);
            if (e != cudaSuccess) return e;
            raised = true;
        }
    }
    gae_tma_kernel<E, HAS_MASK, WANT_PARTIALS, WRITE_RET><<<blocks, E, 2 * STAGE, s>>>(rewards, values, dones, last_values, last_dones, mask,
                                                                                         advs, returns, partials, H, N, g, gt);
    return cudaGetLastError();
}

// tile width: 64 envs per CTA while that still gives at most ~4 CTAs per SM worth of tiles (small N: more SMs pulling), else 128
static inline int gae_tile_envs(int N, int max_blocks) {
    const int b64 = (N + 63) / 64;
    return (N <= 148 * 4 * 64 && (max_blocks <= 0 || b64 <= max_blocks)) ? 64 : 128;
}
static inline bool gae_tma_ok(const void* a, const void* b, const void* c, const void* d, const void* e, const void* f, int N) {
    auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return N % 16 == 0 && al(a) && al(b) && al(c) && al(d) && al(e) && al(f);
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
    if (dense && dones_is_u8 && last_dones_is_u8 && b200rl_gae_use_tma() &&
        gae_tma_ok(rewards, values, dones, advs, returns, nullptr, N)) {
        const uint8_t* d8 = (const uint8_t*)dones; const uint8_t* ld8 = (const uint8_t*)last_dones;
        cudaError_t le;
        if (gae_tile_envs(N, 0) == 64)
            le = returns ? launch_gae_tma<64, false, false, true>(rewards, values, d8, last_values, ld8, nullptr, advs, returns, nullptr, H, N, g, gt, as_stream(stream))
                         : launch_gae_tma<64, false, false, false>(rewards, values, d8, last_values, ld8, nullptr, advs, nullptr, nullptr, H, N, g, gt, as_stream(stream));
        else
            le = returns ? launch_gae_tma<128, false, false, true>(rewards, values, d8, last_values, ld8, nullptr, advs, returns, nullptr, H, N, g, gt, as_stream(stream))
                         : launch_gae_tma<128, false, false, false>(rewards, values, d8, last_values, ld8, nullptr, advs, nullptr, nullptr, H, N, g, gt, as_stream(stream));
        return le == cudaSuccess ? B200RL_OK : (int)le;
    }
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
    cudaStream_t s = as_stream(stream);
    if (b200rl_gae_use_tma() && gae_tma_ok(rewards, values, dones, advs, returns, mask, N)) {
        const int E = gae_tile_envs(N, partials ? max_partials : 0);
        const int nb = (N + E - 1) / E;
        if (n_blocks_out_host) *n_blocks_out_host = nb;
        if (partials && nb > max_partials) return B200RL_EINVAL;
        cudaError_t le;
#define TMA_GO(EE)                                                                                                                          \
    do {                                                                                                                                    \
        if (partials && mask) le = launch_gae_tma<EE, true, true, true>(rewards, values, dones, last_values, last_dones, mask, advs, returns, partials, H, N, g, gt, s); \
        else if (partials) le = launch_gae_tma<EE, false, true, true>(rewards, values, dones, last_values, last_dones, nullptr, advs, returns, partials, H, N, g, gt, s); \
        else le = launch_gae_tma<EE, false, false, true>(rewards, values, dones, last_values, last_dones, nullptr, advs, returns, nullptr, H, N, g, gt, s); \
    } while (0)
        if (E == 64) TMA_GO(64); else TMA_GO(128);
#undef TMA_GO
        return le == cudaSuccess ? B200RL_OK : (int)le;
    }
    const int threads = 128;
    const int blocks = (N + threads - 1) / threads;
    if (n_blocks_out_host) *n_blocks_out_host = blocks;
    if (partials && blocks > max_partials) return B200RL_EINVAL;
    if (partials) {
        if (mask) gae_fused_kernel<uint8_t, uint8_t, true, true, true><<<blocks, threads, 0, s>>>(rewards, values, dones, last_values, last_dones, mask, advs, returns, partials, H, N, g, gt);
        else gae_fused_kernel<uint8_t, uint8_t, false, true, true><<<blocks, threads, 0, s>>>(rewards, values, dones, last_values, last_dones, mask, advs, returns, partials, H, N, g, gt);
    } else {
        gae_fused_kernel<uint8_t, uint8_t, false, false, true><<<blocks, threads, 0, s>>>(rewards, values, dones, last_values, last_dones, nullptr, advs, returns, nullptr, H, N, g, gt);
    }
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}
