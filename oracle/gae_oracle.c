/* C restatement of the reference GAE recursion -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Follows rl_games/triton_kernels/gae_kernel.py:62-79 (_pytorch_gae; same fp32 op order, no FMA
 * contraction: compile with -ffp-contract=off) and tests/test_triton_gae.py:20-42 (fp64 scalar recursion).
 * Used by tests/ as a second, torch-free checker and by bench.py's cpu_baseline GAE micro-baseline.
 * Layout: rewards/values/advs [H][N] (V == 1), dones [H][N] float, last_values [N], last_dones [N]. */
#include <stddef.h>

void gae_oracle_f32(const float* rewards, const float* values, const float* dones, const float* last_values,
                    const float* last_dones, float* advs, int H, int N, double gamma, double tau) {
    const float g = (float)gamma, gt = (float)(gamma * tau);
    for (int e = 0; e < N; ++e) {
        float nv = last_values[e], nnt = 1.0f - last_dones[e], last = 0.0f;
        for (int t = H - 1; t >= 0; --t) {
            const size_t i = (size_t)t * N + e;
            const float t1 = g * nv;
            const float t2 = t1 * nnt;
            const float t3 = rewards[i] + t2;
            const float delta = t3 - values[i];
            const float u1 = gt * nnt;
            const float u2 = u1 * last;
            last = delta + u2;
            advs[i] = last;
            nv = values[i];
            nnt = 1.0f - dones[i];
        }
    }
}

void gae_oracle_f64(const float* rewards, const float* values, const float* dones, const float* last_values,
                    const float* last_dones, double* advs, int H, int N, double gamma, double tau) {
    for (int e = 0; e < N; ++e) {
        double nv = last_values[e], nnt = 1.0 - (double)last_dones[e], last = 0.0;
        for (int t = H - 1; t >= 0; --t) {
            const size_t i = (size_t)t * N + e;
            const double delta = (double)rewards[i] + gamma * nv * nnt - (double)values[i];
            last = delta + gamma * tau * nnt * last;
            advs[i] = last;
            nv = values[i];
            nnt = 1.0 - (double)dones[i];
        }
    }
}
