// Optimiser step: (1/world) scaling + global-norm clip + Adam + on-device adaptive-KL LR schedule.
// Replaces a2c_common.py:493-514 (trancate_gradients_and_step), torch.optim.Adam(eps=1e-8,
// weight_decay, fused=True) (a2c_continuous.py:44-48) and schedulers.py:19-33 + a2c_common.py:1557-1563
// (kl.item() host sync per minibatch).  One launch: every CTA recomputes the global grad norm from the
// (L2-resident) flat gradient buffer in the same order, so no inter-CTA reduction or atomics are needed
// and the result is deterministic; the last CTA to finish advances (step, lr).
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

struct OptCfgDev {
    double beta1, beta2, eps, weight_decay, grad_norm, kl_threshold, min_lr, max_lr, lr_multiplier, grad_scale;
    int truncate_grads, adaptive_lr;
};

struct PackTabDev { int n_seg; int off[4]; int R[4]; int C[4]; unsigned CS[4]; unsigned dst[4]; };

__global__ void __launch_bounds__(1024) adam_step_kernel(float* __restrict__ params, const float* __restrict__ grads,
                                                        float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, int n,
                                                        double* state_d, const float* __restrict__ kl_dev, OptCfgDev c,
                                                        float* __restrict__ stats_out, int* counter, unsigned char* __restrict__ wpack,
                                                        PackTabDev tab) {
    __shared__ double sm[32];
    __shared__ int is_last;
    const double lr = state_d[0];
    const double step = state_d[1] + 1.0;
    // running products beta^step live in the device state (state_d[2], state_d[3]); 0 means "not started" == 1.0
    const double p1 = (state_d[2] > 0.0 ? state_d[2] : 1.0) * c.beta1;
    const double p2 = (state_d[3] > 0.0 ? state_d[3] : 1.0) * c.beta2;
    const float gs = (float)c.grad_scale;
    // ---- global norm (identical in every CTA: same order, fp32 per-thread partials, fp64 block reduction) ----
    double acc[1] = {0.0};
    if (c.truncate_grads || stats_out) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const int n4 = ((reinterpret_cast<uintptr_t>(grads) & 15) == 0) ? (n >> 2) : 0;
        const float4* g4 = reinterpret_cast<const float4*>(grads);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const float4 g = __ldg(g4 + i);
            a0 = fmaf(g.x * gs, g.x * gs, a0); a1 = fmaf(g.y * gs, g.y * gs, a1);
            a2 = fmaf(g.z * gs, g.z * gs, a2); a3 = fmaf(g.w * gs, g.w * gs, a3);
        }
        for (int i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
            const float g = __ldg(grads + i) * gs;
            a0 = fmaf(g, g, a0);
        }
        acc[0] = ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
        block_sum_d<1>(acc, sm);
    }
    const float total_norm = (float)sqrt(acc[0]);
    float coef = 1.0f;
    if (c.truncate_grads) coef = fminf((float)c.grad_norm / (total_norm + 1e-6f), 1.0f);
    const float b1 = (float)c.beta1, b2 = (float)c.beta2;
    const double bc1 = 1.0 - p1, bc2 = 1.0 - p2;
    const float step_size = (float)(lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float eps = (float)c.eps, wd = (float)c.weight_decay;
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int i0 = blockIdx.x * per, i1 = min(i0 + per, n);
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        float g = __ldg(grads + i) * gs;
        if (c.truncate_grads) g *= coef;
        float p = params[i];
        if (wd != 0.f) g = fmaf(wd, p, g);
        float m = exp_avg[i], v = exp_avg_sq[i];
        m = m + (g - m) * (1.0f - b1);                 // lerp (torch fused adam)
        v = b2 * v + (1.0f - b2) * g * g;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        p -= step_size * m / denom;
        params[i] = p; exp_avg[i] = m; exp_avg_sq[i] = v;
        if (wpack) {
            // refresh the bf16 operand copy the tcgen05 kernels consume (INTERLEAVE weight tile: CS = (Rpad/8)*128, RS = 128)
#pragma unroll
            for (int sgi = 0; sgi < 4; ++sgi) {
                if (sgi < tab.n_seg) {
                    const int loc = i - tab.off[sgi];
                    if (loc >= 0 && loc < tab.R[sgi] * tab.C[sgi]) {
                        const int r = loc / tab.C[sgi], cc = loc - r * tab.C[sgi];
                        const unsigned o = tab.dst[sgi] + (unsigned)(r & 7) * 16u + (unsigned)(cc >> 3) * tab.CS[sgi] + (unsigned)(r >> 3) * 128u +
                                           (unsigned)(cc & 7) * 2u;
                        *reinterpret_cast<__nv_bfloat16*>(wpack + o) = __float2bfloat16_rn(p);
                    }
                }
            }
        }
    }
    // ---- last CTA advances step / lr (all CTAs have read them by now) ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int prev = atomicAdd(counter, 1);
        is_last = (prev == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        double new_lr = lr;
        if (c.adaptive_lr && kl_dev) {
            const double kl = (double)(__ldg(kl_dev) * gs);   // summed over ranks by the all-reduce -> mean (a2c_common.py:1559-1561)
            // schedulers.py:19-33 (python floats == fp64)
            if (kl > 2.0 * c.kl_threshold) new_lr = fmax(lr / c.lr_multiplier, c.min_lr);
            if (kl < 0.5 * c.kl_threshold) new_lr = fmin(lr * c.lr_multiplier, c.max_lr);
        }
        state_d[0] = new_lr;
        state_d[1] = step;
        state_d[2] = p1;
        state_d[3] = p2;
        if (stats_out) { stats_out[B200RL_STAT_LR] = (float)lr; stats_out[B200RL_STAT_GNORM] = total_norm; }
        *counter = 0;
    }
}

}  // namespace

B200RL_EXPORT int b200rl_adam_step_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int n,
                                       double* state_d, const float* kl_dev, const b200rl_opt_cfg* cfg_host,
                                       float* stats_out, int* counter, void* wpack, const b200rl_pack_table* tab_host,
                                       void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !state_d || !cfg_host || !counter || n <= 0) return B200RL_EINVAL;
    if ((wpack != nullptr) != (tab_host != nullptr)) return B200RL_EINVAL;
    PackTabDev tab{};
    if (tab_host) {
        if (tab_host->n_seg < 0 || tab_host->n_seg > 4) return B200RL_EINVAL;
        tab.n_seg = tab_host->n_seg;
        for (int i = 0; i < tab.n_seg; ++i) {
            tab.off[i] = tab_host->flat_off[i]; tab.R[i] = tab_host->rows[i]; tab.C[i] = tab_host->cols[i];
            tab.CS[i] = tab_host->cs_bytes[i]; tab.dst[i] = tab_host->dst_off[i];
        }
    }
    OptCfgDev c;
    c.beta1 = cfg_host->beta1; c.beta2 = cfg_host->beta2; c.eps = cfg_host->eps; c.weight_decay = cfg_host->weight_decay;
    c.grad_norm = cfg_host->grad_norm; c.kl_threshold = cfg_host->kl_threshold; c.min_lr = cfg_host->min_lr;
    c.max_lr = cfg_host->max_lr; c.lr_multiplier = cfg_host->lr_multiplier; c.grad_scale = cfg_host->grad_scale;
    c.truncate_grads = cfg_host->truncate_grads; c.adaptive_lr = cfg_host->adaptive_lr;
    int blocks = (n + 2047) / 2048;       // two elements per thread in the update; every CTA re-derives the norm from L2
    if (blocks > 148) blocks = 148;
    if (blocks < 1) blocks = 1;
    adam_step_kernel<<<blocks, 1024, 0, as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, state_d, kl_dev, c, stats_out,
                                                            counter, (unsigned char*)wpack, tab);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}
