// LSTM-with-dones cell kernels (fp32): the pointwise half of common/layers/recurrent.py:20-80 (LSTMWithDones ->
// torch.nn.LSTM, gate order i,f,g,o) and of its BPTT, for the rollout (one step) and the seq_length-window training
// forward/backward.  The GEMM halves (x W_ih^T, h W_hh^T and their dgrad/wgrad) run on the linear_* kernels.
// State zeroing at episode boundaries (a2c_common.py:1150-1153 in the rollout, RnnWithDones.forward :26-58 in training) is
// folded in as a multiplication of the carried (h, c) by (1 - done) when they are handed to the next step.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// gates: [S, 4*Hd] pre-activations in, activations (i,f,g,o) out.   cin: [S,Hd] (already masked c_{t-1})
// outputs: c_out [S,Hd]; h_out [S,Hd] dense; optional h_scatter: row s -> (s / spc) * scatter_stride + s % spc  (MLP-order buffer)
// optional next-step carries: hin_next = h * (1 - done_next[s]), cin_next = c * (1 - done_next[s]) with done_next rows chunk-mapped.
__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(float* __restrict__ gates, const float* __restrict__ cin, float* __restrict__ c_out,
                                                           float* __restrict__ h_out, float* __restrict__ h_scatter, int spc,
                                                           int64_t scatter_stride, float* __restrict__ hin_next, float* __restrict__ cin_next,
                                                           const uint8_t* __restrict__ done_next, int d_rpc, int64_t d_stride, int S, int Hd) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)S * Hd) return;
    const int s = (int)(idx / Hd), k = (int)(idx - (int64_t)s * Hd);
    float* g = gates + (int64_t)s * 4 * Hd;
    const float i = sigmoidf_(g[k]), f = sigmoidf_(g[Hd + k]), gg = tanhf(g[2 * Hd + k]), o = sigmoidf_(g[3 * Hd + k]);
    g[k] = i; g[Hd + k] = f; g[2 * Hd + k] = gg; g[3 * Hd + k] = o;
    const float c = f * cin[idx] + i * gg;
    const float h = o * tanhf(c);
    c_out[idx] = c;
    h_out[idx] = h;
    if (h_scatter) h_scatter[((int64_t)(s / spc) * scatter_stride + (s % spc)) * Hd + k] = h;
    if (hin_next) {
        float m = 1.0f;
        if (done_next) m = 1.0f - (float)done_next[chunk_row(s, d_rpc, d_stride)];
        hin_next[idx] = h * m;
        cin_next[idx] = c * m;
    }
}

// Backward of one cell step.  dh_in = dH_mlp[row(s)] (chunk scatter order, optional) + dhin_next[s] * m_next ; dc_in = dcin_next * m_next.
// Writes dgates [S,4Hd] (pre-activation gradients) and dcin [S,Hd] (gradient w.r.t. the masked c input of THIS step).
__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(const float* __restrict__ gates_act, const float* __restrict__ c_t,
                                                           const float* __restrict__ cin, const float* __restrict__ dH, int spc,
                                                           int64_t scatter_stride, const float* __restrict__ dhin_next,
                                                           const float* __restrict__ dcin_next, const uint8_t* __restrict__ done_next, int d_rpc,
                                                           int64_t d_stride, float* __restrict__ dgates, float* __restrict__ dcin, int S, int Hd) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)S * Hd) return;
    const int s = (int)(idx / Hd), k = (int)(idx - (int64_t)s * Hd);
    const float* g = gates_act + (int64_t)s * 4 * Hd;
    const float i = g[k], f = g[Hd + k], gg = g[2 * Hd + k], o = g[3 * Hd + k];
    float m = 1.0f;
    if (done_next) m = 1.0f - (float)done_next[chunk_row(s, d_rpc, d_stride)];
    float dh = dH ? dH[((int64_t)(s / spc) * scatter_stride + (s % spc)) * Hd + k] : 0.f;
    float dc = 0.f;
    if (dhin_next) { dh += dhin_next[idx] * m; dc += dcin_next[idx] * m; }
    const float tc = tanhf(c_t[idx]);
    dc += dh * o * (1.0f - tc * tc);
    float* dg = dgates + (int64_t)s * 4 * Hd;
    dg[k] = dc * gg * i * (1.0f - i);
    dg[Hd + k] = dc * cin[idx] * f * (1.0f - f);
    dg[2 * Hd + k] = dc * i * (1.0f - gg * gg);
    dg[3 * Hd + k] = dh * tc * o * (1.0f - o);
    dcin[idx] = dc * f;
}

// out[s][k] = in[row(s)][k] * (1 - done[drow(s)])   (masked initial state of a window / zeroing after an episode end)
__global__ void __launch_bounds__(256) rnn_mask_rows_kernel(const float* __restrict__ in, int in_rpc, int64_t in_stride, float* __restrict__ out,
                                                           const uint8_t* __restrict__ done, int d_rpc, int64_t d_stride, int S, int Hd) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)S * Hd) return;
    const int s = (int)(idx / Hd), k = (int)(idx - (int64_t)s * Hd);
    const float m = done ? 1.0f - (float)done[chunk_row(s, d_rpc, d_stride)] : 1.0f;
    out[idx] = in[chunk_row(s, in_rpc, in_stride) * Hd + k] * m;
}

// Train-time reset flags of an RNN policy on a next_step-autoreset env (a2c_common.py:1180-1191): besides the buffer's own dones the
// state reset also fires ENTERING the first real row after a filler reset row (valid == 0), mirroring the rollout-side re-zero.
//   out[t, n] = dones[t, n] | (t > 0 && valid[t - 1, n] == 0)
__host__ __device__ inline uint8_t rnn_train_done(const uint8_t* dones, const float* valid, int64_t i, int N) {
    return (uint8_t)((dones[i] != 0) || (i >= N && valid[i - N] == 0.0f));
}
__global__ void __launch_bounds__(256) rnn_train_dones_kernel(const uint8_t* __restrict__ dones, const float* __restrict__ valid,
                                                              uint8_t* __restrict__ out, int64_t n, int N) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = rnn_train_done(dones, valid, i, N);
}

}  // namespace

B200RL_EXPORT int b200rl_lstm_cell_fwd_f32(float* gates, const float* cin, float* c_out, float* h_out, float* h_scatter, int scatter_rpc,
                                           int64_t scatter_stride, float* hin_next, float* cin_next, const uint8_t* done_next,
                                           int done_rpc, int64_t done_stride, int S, int Hd, void* stream) {
    if (!gates || !cin || !c_out || !h_out || S <= 0 || Hd <= 0) return B200RL_EINVAL;
    if ((hin_next == nullptr) != (cin_next == nullptr)) return B200RL_EINVAL;
    if (h_scatter && scatter_rpc <= 0) return B200RL_EINVAL;
    if (done_next && done_rpc <= 0) return B200RL_EINVAL;
    const int64_t n = (int64_t)S * Hd;
    lstm_cell_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(gates, cin, c_out, h_out, h_scatter, scatter_rpc > 0 ? scatter_rpc : 1,
                                                                                     scatter_stride, hin_next, cin_next, done_next,
                                                                                     done_rpc > 0 ? done_rpc : 1, done_stride, S, Hd);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}

B200RL_EXPORT int b200rl_lstm_cell_bwd_f32(const float* gates_act, const float* c_t, const float* cin, const float* dH, int scatter_rpc,
                                           int64_t scatter_stride, const float* dhin_next, const float* dcin_next, const uint8_t* done_next,
                                           int done_rpc, int64_t done_stride, float* dgates, float* dcin, int S, int Hd, void* stream) {
    if (!gates_act || !c_t || !cin || !dgates || !dcin || S <= 0 || Hd <= 0) return B200RL_EINVAL;
    if ((dhin_next == nullptr) != (dcin_next == nullptr)) return B200RL_EINVAL;
    const int64_t n = (int64_t)S * Hd;
    lstm_cell_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(gates_act, c_t, cin, dH, scatter_rpc > 0 ? scatter_rpc : 1,
                                                                                     scatter_stride, dhin_next, dcin_next, done_next,
                                                                                     done_rpc > 0 ? done_rpc : 1, done_stride, dgates, dcin, S, Hd);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}

B200RL_EXPORT int b200rl_rnn_mask_rows_f32(const float* in, int in_rpc, int64_t in_stride, float* out, const uint8_t* done, int done_rpc,
                                           int64_t done_stride, int S, int Hd, void* stream) {
    if (!in || !out || S <= 0 || Hd <= 0 || in_rpc <= 0) return B200RL_EINVAL;
    const int64_t n = (int64_t)S * Hd;
    rnn_mask_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(in, in_rpc, in_stride, out, done, done_rpc > 0 ? done_rpc : 1,
                                                                                     done_stride, S, Hd);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}

B200RL_EXPORT int b200rl_rnn_train_dones_u8(const uint8_t* dones, const float* valid, uint8_t* out, int H, int N, void* stream) {
    if (!dones || !valid || !out || H <= 0 || N <= 0) return B200RL_EINVAL;
    const int64_t n = (int64_t)H * N;
    rnn_train_dones_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(dones, valid, out, n, N);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}

#ifdef B200RL_TEST_HOOKS   // test-only host entry points: compiled into tests/libb200rl_testhooks.so (csrc/build.py), not into the product library
// host test entry point (tests/test_rnn_rows_cpu.py): the element function above over HOST arrays; not in include/b200rl.h
B200RL_EXPORT int b200rl_hosttest_rnn_train_dones(const uint8_t* dones, const float* valid, uint8_t* out, int H, int N) {
    for (int64_t i = 0; i < (int64_t)H * N; ++i) out[i] = rnn_train_done(dones, valid, i, N);
    return B200RL_OK;
}
#endif  // B200RL_TEST_HOOKS
