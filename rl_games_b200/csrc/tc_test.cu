// Bring-up / regression kernel for the tcgen05 building blocks (used by tests/test_tc_gpu.py):
// D[128, N] = A[128, K] . B[N, K]^T with bf16 operands staged by hand into INTERLEAVE smem tiles, fp32 TMEM
// accumulators read back with tcgen05.ld.  a_mn / b_mn select MN-major operand views: the operand is then
// supplied transposed in global memory (A^T: [K,128], B^T: [K,N]) and staged as a [K rows x MN cols] tile.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

__global__ void __launch_bounds__(128) tc_gemm_test_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                          float* __restrict__ D, int N, int K, int a_mn, int b_mn) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int M = 128;
    // tile geometry: rows x cols with cols contiguous in 16B chunks
    const int a_rows = a_mn ? K : M, a_cols = a_mn ? M : K;
    const int b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
    const uint32_t a_CS = 128, a_RS = (uint32_t)(a_cols / 8) * 128;     // column groups adjacent, then row groups
    const uint32_t b_CS = 128, b_RS = (uint32_t)(b_cols / 8) * 128;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)a_rows * a_cols * 2;
    // stage (global row-major [rows, cols] bf16 -> interleave tile)
    for (int i = tid; i < a_rows * (a_cols / 8); i += blockDim.x) {
        const int r = i / (a_cols / 8), cg = i % (a_cols / 8);
        *reinterpret_cast<uint4*>(sA + tc::tile_off(r, cg, a_CS, a_RS)) = *reinterpret_cast<const uint4*>(A + (size_t)r * a_cols + cg * 8);
    }
    for (int i = tid; i < b_rows * (b_cols / 8); i += blockDim.x) {
        const int r = i / (b_cols / 8), cg = i % (b_cols / 8);
        *reinterpret_cast<uint4*>(sB + tc::tile_off(r, cg, b_CS, b_RS)) = *reinterpret_cast<const uint4*>(B + (size_t)r * b_cols + cg * 8);
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(M, N, a_mn, b_mn);
        for (int k = 0; k < K / 16; ++k) {
            // K-major: one K=16 step = 2 column groups -> +2*CS ; LBO = CS, SBO = RS
            // MN-major: one K=16 step = 2 row groups   -> +2*RS ; LBO = RS (K stride), SBO = CS (MN stride)
            const uint64_t ad = a_mn ? tc::make_smem_desc(tc::smem_u32(sA) + k * 2 * a_RS, a_RS, a_CS)
                                     : tc::make_smem_desc(tc::smem_u32(sA) + k * 2 * a_CS, a_CS, a_RS);
            const uint64_t bd = b_mn ? tc::make_smem_desc(tc::smem_u32(sB) + k * 2 * b_RS, b_RS, b_CS)
                                     : tc::make_smem_desc(tc::smem_u32(sB) + k * 2 * b_CS, b_CS, b_RS);
            tc::umma_bf16(tmem, ad, bd, idesc, k > 0);
        }
        tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace

B200RL_EXPORT int b200rl_tc_gemm_test(const void* A_bf16, const void* B_bf16, float* D, int N, int K, int a_mn, int b_mn,
                                      void* stream) {
    if (!A_bf16 || !B_bf16 || !D || N < 16 || N > 256 || N % 16 || K < 16 || K % 16) return B200RL_EINVAL;
    const size_t smem = (size_t)128 * K * 2 + (size_t)N * K * 2;
    if (smem > 200 * 1024) return B200RL_EINVAL;
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    tc_gemm_test_kernel<<<1, 128, smem, as_stream(stream)>>>((const __nv_bfloat16*)A_bf16, (const __nv_bfloat16*)B_bf16, D, N, K,
                                                            a_mn, b_mn);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}
