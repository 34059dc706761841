// tcgen05 / TMEM / mbarrier PTX wrappers for sm_100a (5th-gen tensor cores).  Hand-written equivalents of the
// CUTLASS sm100 primitives; descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp.
//
// Shared-memory operand tiles use the canonical INTERLEAVE (no-swizzle) layout: 16-byte chunks of 8 bf16
// that are contiguous along the tile's "column" index, 8 consecutive "rows" stacked at 16 B stride (one 128 B
// core matrix), with two free strides:
//     offset(r, c) = (r % 8) * 16 + (c % 8) * 2 + (c / 8) * CS + (r / 8) * RS            [bytes]
// The SAME bytes serve as a K-major operand (rows = M/N, cols = K: LBO = CS, SBO = RS) and as an MN-major
// operand (cols = M/N, rows = K: SBO = CS, LBO = RS) -- which is what lets one activation / delta tile feed
// the forward/dgrad MMAs (K-major) and the weight-gradient MMAs (MN-major) without any transposed copy.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- descriptors -------------------------------------------------------------------------------
// SmemDescriptor: start_address[0,14) (>>4), leading_byte_offset[16,30) (>>4), stride_byte_offset[32,46) (>>4),
// version[46,48) = 1 (Blackwell), base_offset[49,52) = 0, lbo_mode[52] = 0, layout_type[61,64) = 0 (INTERLEAVE)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// InstrDescriptor (kind::f16): c_format[4,6)=1 (f32), a_format[7,10)=1 (bf16), b_format[10,13)=1 (bf16),
// a_major[15], b_major[16] (0 = K-major, 1 = MN-major), n_dim[17,23) = N>>3, m_dim[24,29) = M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM --------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp, .sync.aligned
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// tcgen05.commit: arrive on the mbarrier when all previously issued MMAs of this thread have completed
// (implicitly performs tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- MMA: D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate; issued by ONE thread ----
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32-bit, 16 consecutive columns per call (warp w%4 owns lanes 32*(w%4)..+31)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// pack 8 floats -> 8 bf16 (16 bytes)
__host__ __device__ __forceinline__ uint4 pack8_bf16(const float* f) {
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    return u;
}

// byte offset of the 16-byte chunk holding (row r, columns 8*cg .. 8*cg+7) in an INTERLEAVE tile
__host__ __device__ __forceinline__ uint32_t tile_off(int r, int cg, uint32_t CS, uint32_t RS) {
    return (uint32_t)(r & 7) * 16u + (uint32_t)cg * CS + (uint32_t)(r >> 3) * RS;
}

}  // namespace tc
