// fp32 (CUDA-core) MLP building blocks: the `mixed_precision: False` path and the tight-parity
// reference for the tcgen05 bf16 path (mlp_tc.cu).  network_builder.py:494-512 forward and its autograd.
//   fwd : Y[M,N]  = act( norm(X)[M,K] . W[N,K]^T + b )
//   dgrad: dX[M,K] = ( dY[M,N] . W[N,K] ) * act'(A_prev)
//   wgrad: dWp[s][N,K] = dY_s^T . norm(X_s) ; dbp[s][N] = colsum(dY_s)     (rows split s = blockIdx.z)
// Register-tiled SGEMM: 128x64 block tile, 8x4 per thread, BK=16, operands staged in shared memory.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;

__device__ __forceinline__ float load_x(const float* __restrict__ X, int64_t row_off, int k, int K,
                                        const float* __restrict__ nm, const float* __restrict__ ns) {
    if (k >= K) return 0.f;
    float v = __ldg(X + row_off + k);
    if (nm) {
        v = __fdiv_rn(__fsub_rn(v, __ldg(nm + k)), __ldg(ns + k));
        v = fminf(fmaxf(v, -5.0f), 5.0f);
    }
    return v;
}

// ---------------- forward ----------------
__global__ void __launch_bounds__(NT) linear_fwd_kernel(
    const float* __restrict__ X, int rows_per_chunk, int64_t chunk_stride, int64_t x_ld,
    const float* __restrict__ nm, const float* __restrict__ ns,
    const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ Y,
    int M, int K, int N, int act, int accumulate) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int ty = tid / 16, tx = tid % 16;   // 16 x 16 threads; thread tile 8 (M) x 4 (N)
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // loader mapping: A: row = tid/2 (0..127), kk = (tid%2)*8 .. +8 ; B: row = tid/4 (0..63), kk = (tid%4)*4 .. +4
    const int a_row = tid >> 1, a_k = (tid & 1) * 8;
    const int b_row = tid >> 2, b_k = (tid & 3) * 4;
    const int gm = m0 + a_row;
    const int64_t a_off = gm < M ? chunk_row(gm, rows_per_chunk, chunk_stride) * x_ld : 0;
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = k0 + a_k + i;
            As[a_k + i][a_row] = (gm < M) ? load_x(X, a_off, k, K, nm, ns) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + b_k + i, n = n0 + b_row;
            Bs[b_k + i][b_row] = (n < N && k < K) ? __ldg(W + (int64_t)n * K + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[4];
            *(float4*)&a[0] = *(const float4*)&As[k][ty * 8];
            *(float4*)&a[4] = *(const float4*)&As[k][ty * 8 + 4];
            *(float4*)&b[0] = *(const float4*)&Bs[k][tx * 4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) {
                const float prev = accumulate ? Y[(int64_t)m * N + n] : 0.f;
                Y[(int64_t)m * N + n] = act_fwd(acc[i][j] + (bias ? __ldg(bias + n) : 0.f) + prev, act);
            }
        }
    }
}

// ---------------- dgrad: dX[M,K] = (dY[M,N] . W[N,K]) * act'(A_prev[M,K]) ----------------
__global__ void __launch_bounds__(NT) linear_dgrad_kernel(
    const float* __restrict__ dY, const float* __restrict__ W, const float* __restrict__ A_prev,
    float* __restrict__ dX, int M, int K, int N, int act_prev) {
    __shared__ float As[BK][BM + 4];   // dY tile, reduction dim n
    __shared__ float Bs[BK][BN + 4];   // W tile [n][k]
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, k0o = blockIdx.y * BN;   // output cols = K dimension
    const int ty = tid / 16, tx = tid % 16;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int a_row = tid >> 1, a_n = (tid & 1) * 8;
    const int b_n = tid >> 4, b_k = (tid & 15) * 4;    // 16 n-rows x 64 k
    const int gm = m0 + a_row;
    for (int n0 = 0; n0 < N; n0 += BK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = n0 + a_n + i;
            As[a_n + i][a_row] = (gm < M && n < N) ? __ldg(dY + (int64_t)gm * N + n) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = n0 + b_n, k = k0o + b_k + i;
            Bs[b_n][b_k + i] = (n < N && k < K) ? __ldg(W + (int64_t)n * K + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[4];
            *(float4*)&a[0] = *(const float4*)&As[k][ty * 8];
            *(float4*)&a[4] = *(const float4*)&As[k][ty * 8 + 4];
            *(float4*)&b[0] = *(const float4*)&Bs[k][tx * 4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0o + tx * 4 + j;
            if (k < K) {
                const float d = A_prev ? act_bwd_from_out(__ldg(A_prev + (int64_t)m * K + k), act_prev) : 1.f;
                dX[(int64_t)m * K + k] = acc[i][j] * d;
            }
        }
    }
}

// ---------------- wgrad: dWp[s][N,K] = dY_s^T . X_s ; dbp[s][N] ----------------
// block tile 64 (n) x 64 (k); thread tile 4x4; reduction over rows of split s in slabs of 16 rows.
__global__ void __launch_bounds__(NT) linear_wgrad_kernel(
    const float* __restrict__ dY, const float* __restrict__ X, int rows_per_chunk, int64_t chunk_stride, int64_t x_ld,
    const float* __restrict__ nm, const float* __restrict__ ns,
    float* __restrict__ dWp, float* __restrict__ dbp, int M, int K, int N, int rows_per_split, int64_t split_stride) {
    __shared__ float As[BK][64 + 4];   // dY slab [m][n]
    __shared__ float Bs[BK][64 + 4];   // X slab  [m][k]
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64, s = blockIdx.z;
    const int ty = tid / 16, tx = tid % 16;   // thread tile: n = ty*4.., k = tx*4..
    const int mbeg = s * rows_per_split, mend = min(mbeg + rows_per_split, M);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;   // threads 0..63 accumulate db for column n0+tid when blockIdx.y == 0
    const int l_m = tid >> 4, l_c = (tid & 15) * 4;   // 16 rows x 64 cols, 4 per thread
    for (int mm = mbeg; mm < mend; mm += BK) {
        const int m = mm + l_m;
        const bool mv = m < mend;
        const int64_t xoff = mv ? chunk_row(m, rows_per_chunk, chunk_stride) * x_ld : 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = n0 + l_c + i, k = k0 + l_c + i;
            As[l_m][l_c + i] = (mv && n < N) ? __ldg(dY + (int64_t)m * N + n) : 0.f;
            Bs[l_m][l_c + i] = mv ? load_x(X, xoff, k, K, nm, ns) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < BK; ++r) {
            float a[4], b[4];
            *(float4*)&a[0] = *(const float4*)&As[r][ty * 4];
            *(float4*)&b[0] = *(const float4*)&Bs[r][tx * 4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (blockIdx.y == 0 && tid < 64) {
#pragma unroll
            for (int r = 0; r < BK; ++r) bsum += As[r][tid];
        }
        __syncthreads();
    }
    float* out = dWp + (int64_t)s * split_stride;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < K) out[(int64_t)n * K + k] = acc[i][j];
        }
    }
    if (dbp && blockIdx.y == 0 && tid < 64 && n0 + tid < N) dbp[(int64_t)s * split_stride + n0 + tid] = bsum;
}

__global__ void reduce_splits_kernel(const float* __restrict__ part, float* __restrict__ out, int n, int n_splits,
                                     int64_t split_stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < n_splits; ++k) s += __ldg(part + (int64_t)k * split_stride + i);
    out[i] = s;
}

}  // namespace

B200RL_EXPORT int b200rl_linear_fwd_f32(const float* X, int rows_per_chunk, int64_t chunk_stride, int64_t x_ld,
                                        const float* norm_mean, const float* norm_std,
                                        const float* W, const float* b, float* Y, int M, int K, int Nout, int act,
                                        int accumulate, void* stream) {
    if (!X || !W || !Y || M <= 0 || K <= 0 || Nout <= 0 || rows_per_chunk <= 0) return B200RL_EINVAL;
    if ((norm_mean == nullptr) != (norm_std == nullptr)) return B200RL_EINVAL;
    dim3 grid((M + BM - 1) / BM, (Nout + BN - 1) / BN);
    linear_fwd_kernel<<<grid, NT, 0, as_stream(stream)>>>(X, rows_per_chunk, chunk_stride, x_ld, norm_mean, norm_std, W, b, Y,
                                                          M, K, Nout, act, accumulate);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}

B200RL_EXPORT int b200rl_linear_bwd_data_f32(const float* dY, const float* W, const float* A_prev, float* dX,
                                             int M, int K, int Nout, int act_prev, void* stream) {
    if (!dY || !W || !dX || M <= 0 || K <= 0 || Nout <= 0) return B200RL_EINVAL;
    dim3 grid((M + BM - 1) / BM, (K + BN - 1) / BN);
    linear_dgrad_kernel<<<grid, NT, 0, as_stream(stream)>>>(dY, W, A_prev, dX, M, K, Nout, act_prev);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}

B200RL_EXPORT int b200rl_linear_bwd_weight_f32(const float* dY, const float* X, int rows_per_chunk, int64_t chunk_stride,
                                               int64_t x_ld, const float* norm_mean, const float* norm_std,
                                               float* dW_part, float* db_part, int64_t split_stride, int M, int K, int Nout,
                                               int n_splits, void* stream) {
    if (!dY || !X || !dW_part || M <= 0 || K <= 0 || Nout <= 0 || n_splits <= 0 || rows_per_chunk <= 0) return B200RL_EINVAL;
    if ((norm_mean == nullptr) != (norm_std == nullptr)) return B200RL_EINVAL;
    int rps = (M + n_splits - 1) / n_splits;
    rps = ((rps + BK - 1) / BK) * BK;
    if ((int64_t)rps * (n_splits - 1) >= M && n_splits > 1) {
        // trailing splits would be empty: they still write zeros, which is what the reducer expects
    }
    dim3 grid((Nout + 63) / 64, (K + 63) / 64, n_splits);
    linear_wgrad_kernel<<<grid, NT, 0, as_stream(stream)>>>(dY, X, rows_per_chunk, chunk_stride, x_ld, norm_mean, norm_std,
                                                            dW_part, db_part, M, K, Nout, rps, split_stride);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}

B200RL_EXPORT int b200rl_reduce_splits_f32(const float* part, float* out, int n, int n_splits, int64_t split_stride,
                                           void* stream) {
    if (!part || !out || n <= 0 || n_splits <= 0) return B200RL_EINVAL;
    reduce_splits_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(part, out, n, n_splits, split_stride);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}
