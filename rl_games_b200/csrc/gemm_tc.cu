// Layer-wise bf16 tensor-core GEMMs (tcgen05 / TMEM) for ARBITRARY layer widths: the `mixed_precision: True` path of everything the
// fused MLP kernels (mlp_tc.cu) have no geometry for -- LSTM gate GEMMs (common/layers/recurrent.py:20-80, network_builder.py:452-492:
// [S, in + hidden] x [4 hidden] per step), MLPs wider / deeper than [256,128,64] (configs/mujoco/humanoid_envpool.yaml:23
// [512,256,128]).  Same three operations, same signatures and same fp32 row-major operands as the CUDA-core building blocks of
// mlp_simt.cu, so the host code composes them identically; only the products run on the tensor cores:
//   fwd   : Y[M,N]   = act( norm(X)[M,K] . W[N,K]^T + b [+ Y] )
//   dgrad : dX[M,K]  = ( dY[M,N] . W[N,K] ) * act'(A_prev)
//   wgrad : dW_s[N,K] = dY_s^T . norm(X_s),  db_s[N] = colsum(dY_s)          (rows split s = blockIdx.z, summed by reduce_splits)
// Precision contract = the reference's bf16 autocast of nn.Linear / nn.LSTM (a2c_continuous.py:173): operands rounded to bf16 while
// they are staged into shared memory, fp32 accumulation in TMEM, fp32 bias / activation / outputs.
//
// One CTA = one [128 x BN] output tile (BN <= 256 TMEM columns), reduction in 64-element slabs through a two-stage shared-memory ring.
// Data movement: all 256 threads load one slab of BOTH operands into registers (16-byte vector loads, every load issued before the
// first use: one global round trip per slab), convert to bf16 and store it as INTERLEAVE operand tiles, then issue the loads of the NEXT
// slab before the barrier -- they are in flight while the single issuing thread's MMAs run and while the next stage is awaited; a stage
// is reused when the tcgen05.commit of its MMAs has arrived.  Two CTAs per SM alternate, so the tensor pipe and the memory system stay
// busy while either waits.  Operand views (tc_common.cuh): K-major when the source's contiguous index is the reduction index (X and W
// in fwd, dY in dgrad), MN-major when it is the output index (W in dgrad, dY and X in wgrad) -- no transposes anywhere.  Epilogue: each
// warp transposes its 32 x 32 accumulator patches through shared memory so that every global access (bias, accumulate, activation
// derivative input, output) is a whole 128-byte row segment.  These GEMMs keep fp32 activations in HBM on both sides (the CUDA-core
// path's contract), so at K <= 512 they are bound by HBM / L2 bandwidth, not by the tensor pipe: tools/gemm_tc_sweep.py prints both bounds.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int GT = 256;              // threads per CTA: lane quarter q = warp & 3, column half h = warp >> 2
constexpr int SLAB = 64;             // reduction elements per ring stage
constexpr int BM = 128;              // output rows per CTA (TMEM lanes)
constexpr int BNMAX = 256;           // output columns per CTA (TMEM columns)
constexpr uint32_t A_BYTES = BM * SLAB * 2, B_BYTES = BNMAX * SLAB * 2, STAGE_BYTES = A_BYTES + B_BYTES;

// fp32 source matrix [rows, cols]; rows may be chunked (arena addressing) and normalised on the fly (first layer reads observations)
struct Src {
    const float* p; int64_t ld; int rows_per_chunk; int64_t chunk_stride; const float* nm; const float* ns;
    int rows, cols;          // logical extent (everything outside reads as 0)
    const __nv_bfloat16* pb; // optional bf16 copy of a DENSE source (weights: refreshed once per optimiser step by b200rl_cast_bf16):
                             // staged with plain 16 / 8-byte copies, no conversion, half the bytes
};
__device__ __forceinline__ const float* src_row(const Src& s, int r) {
    return s.p + (s.rows_per_chunk > 0 ? chunk_row(r, s.rows_per_chunk, s.chunk_stride) : (int64_t)r) * s.ld;
}

// ---- staging: global -> registers -> bf16 INTERLEAVE operand tile ------------------------------------------------------------------
// A tile is [tr rows x ncg * 8 cols] of 16-byte bf16 chunks (CS = 128, RS = ncg * 128; tc_common.cuh); tr is a multiple of 8 and ncg a
// power of two (8, 16 or 32).  Item i of a tile = chunk (r, cg) with  r = (i & 7) | ((i >> (3 + lg ncg)) << 3),  cg = (i >> 3) & (ncg - 1):
// the 32 lanes of a warp cover 8 rows x 4 adjacent chunks, so a warp's global loads are 8 full 128-byte lines (fp32) and its 16-byte
// shared-memory stores fall into 4 conflict-free wavefronts (8 consecutive lanes write one 128-byte core matrix).
// Loading and storing are separate so that EVERY load of a slab (both operands) is in flight before the first use, and so that the
// loads of slab s + 1 can be issued before the barrier / MMA issue of slab s (register prefetch).
struct TileGeo { int r0, c0, lg, n_items, row_end; };     // tile origin in the source, log2(ncg), tr * ncg, rows >= row_end read as 0
__device__ __forceinline__ void item_rc(int i, int lg, int& r, int& cg) {
    cg = (i >> 3) & ((1 << lg) - 1);
    r = (i & 7) | ((i >> (3 + lg)) << 3);
}

template <int ITEMS>
__device__ __forceinline__ void load_f32(float4 (&va)[ITEMS], float4 (&vb)[ITEMS], const Src& s, const TileGeo& g, int tid) {
    const bool vec = ((s.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(s.p) & 15) == 0) && ((g.c0 & 3) == 0);
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int i = tid + it * GT;
        va[it] = make_float4(0.f, 0.f, 0.f, 0.f); vb[it] = va[it];
        if (i < g.n_items) {
            int r, cg; item_rc(i, g.lg, r, cg);
            const int gr = g.r0 + r, gc = g.c0 + cg * 8;
            if (gr < g.row_end && gc < s.cols) {
                const float* src = src_row(s, gr) + gc;
                if (vec && gc + 8 <= s.cols) {
                    va[it] = __ldg(reinterpret_cast<const float4*>(src));
                    vb[it] = __ldg(reinterpret_cast<const float4*>(src) + 1);
                } else {
                    float t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t[j] = (gc + j < s.cols) ? __ldg(src + j) : 0.f;
                    va[it] = make_float4(t[0], t[1], t[2], t[3]); vb[it] = make_float4(t[4], t[5], t[6], t[7]);
                }
            }
        }
    }
}
template <int ITEMS>
__device__ __forceinline__ void store_f32(uint8_t* sT, const float4 (&va)[ITEMS], const float4 (&vb)[ITEMS], const Src& s, const TileGeo& g, int tid) {
    const uint32_t RS = 128u << g.lg;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int i = tid + it * GT;
        if (i < g.n_items) {
            int r, cg; item_rc(i, g.lg, r, cg);
            float f[8] = {va[it].x, va[it].y, va[it].z, va[it].w, vb[it].x, vb[it].y, vb[it].z, vb[it].w};
            if (s.nm) {
                const int gr = g.r0 + r, gc = g.c0 + cg * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    f[j] = (gr < g.row_end && gc + j < s.cols)
                               ? fminf(fmaxf(__fdiv_rn(__fsub_rn(f[j], __ldg(s.nm + gc + j)), __ldg(s.ns + gc + j)), -5.0f), 5.0f) : 0.f;
            }
            *reinterpret_cast<uint4*>(sT + tile_off(r, cg, 128u, RS)) = pack8_bf16(f);
        }
    }
}
// bf16 twin of a dense source (weights): plain 16 / 8-byte copies, no conversion, half the bytes
template <int ITEMS>
__device__ __forceinline__ void load_b16(uint4 (&u)[ITEMS], const Src& s, const TileGeo& g, int tid) {
    const int al = (int)(((s.ld * 2) | (int64_t)(reinterpret_cast<uintptr_t>(s.pb) & 15)) & 15);     // 0: rows 16-byte aligned, 8: 8-byte aligned
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int i = tid + it * GT;
        u[it] = make_uint4(0, 0, 0, 0);
        if (i < g.n_items) {
            int r, cg; item_rc(i, g.lg, r, cg);
            const int gr = g.r0 + r, gc = g.c0 + cg * 8;
            if (gr < g.row_end && gc < s.cols) {
                const __nv_bfloat16* src = s.pb + (int64_t)gr * s.ld + gc;
                if (gc + 8 <= s.cols && al == 0) {
                    u[it] = __ldg(reinterpret_cast<const uint4*>(src));
                } else if (gc + 8 <= s.cols && (al & 7) == 0) {
                    const uint2 a = __ldg(reinterpret_cast<const uint2*>(src)), b = __ldg(reinterpret_cast<const uint2*>(src) + 1);
                    u[it] = make_uint4(a.x, a.y, b.x, b.y);
                } else {
                    __nv_bfloat16 t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t[j] = (gc + j < s.cols) ? src[j] : __float2bfloat16_rn(0.f);
                    u[it] = *reinterpret_cast<const uint4*>(t);
                }
            }
        }
    }
}
template <int ITEMS>
__device__ __forceinline__ void store_b16(uint8_t* sT, const uint4 (&u)[ITEMS], const TileGeo& g, int tid) {
    const uint32_t RS = 128u << g.lg;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int i = tid + it * GT;
        if (i < g.n_items) {
            int r, cg; item_rc(i, g.lg, r, cg);
            *reinterpret_cast<uint4*>(sT + tile_off(r, cg, 128u, RS)) = u[it];
        }
    }
}
__device__ __forceinline__ int ilog2_pow2(int v) { return 31 - __clz(v); }

// forward epilogue rows of one 32 x 32 patch: lane = column, k = row; ACT is a compile-time constant so the loop body is straight-line.
// ELU uses the fast exponential like the fused kernels of mlp_tc.cu (mixed-precision contract: |error| ~ 1e-7, far below the bf16
// operand rounding); tanh stays tanhf.
template <int ACT>
__device__ __forceinline__ void fwd_rows(float* out, const float* sT, const float (&pre)[32], bool has_pre, float bias, int lane, int n_rows,
                                         int ld) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        if (k < n_rows) {
            float x = sT[k * 33 + lane] + bias;
            if (has_pre) x += pre[k];
            if (ACT == B200RL_ACT_ELU) x = x > 0.f ? x : (__expf(x) - 1.0f);
            else if (ACT == B200RL_ACT_RELU) x = x > 0.f ? x : 0.f;
            else if (ACT == B200RL_ACT_TANH) x = tanhf(x);
            out[(int64_t)k * ld] = x;
        }
    }
}

template <int ACT>
__device__ __forceinline__ void dgrad_rows(float* out, const float* sT, const float (&pre)[32], int lane, int n_rows, int ld) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        if (k < n_rows) {
            float x = sT[k * 33 + lane];
            if (ACT == B200RL_ACT_ELU) x *= pre[k] > 0.f ? 1.f : pre[k] + 1.f;          // derivative through the activation OUTPUT (common.cuh)
            else if (ACT == B200RL_ACT_RELU) x = pre[k] > 0.f ? x : 0.f;
            else if (ACT == B200RL_ACT_TANH) x *= 1.f - pre[k] * pre[k];
            out[(int64_t)k * ld] = x;
        }
    }
}

enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

struct GemmArgs {
    Src a, b;                    // the two operand sources (see the mode table in the kernel)
    int Mo, No, R;               // output rows, output cols, reduction length
    int BN;                      // output columns per CTA: 64, 128 or 256
    // epilogue
    const float* bias; float* Y; int act; int accumulate;              // fwd: Y[Mo, No]
    const float* A_prev; float* dX; int act_prev;                       // dgrad: dX[Mo, No]
    float* dW; float* db; int64_t split_stride; int rows_per_split;     // wgrad: dW[s][Mo (= layer N), No (= layer K)], db[s][Mo]
};

// mode        output rows (TMEM lanes)   output cols        reduction   A operand (rows m)            B operand (rows n)
// FWD         batch rows                 layer outputs N    K           X[m, k]   K-major            W[n, k]   K-major
// DGRAD       batch rows                 layer inputs K     N           dY[m, n]  K-major            W[n, k]   MN-major (rows = reduction n)
// WGRAD       layer outputs N            layer inputs K     batch rows  dY[r, n]  MN-major           X[r, k]   MN-major (rows = reduction r)
template <int MODE, bool BB16>
__global__ void __launch_bounds__(GT, 2) gemm_tc_kernel(const GemmArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, h = warp >> 2;
    const int BN = p.BN;                 // 64, 128 or (bf16 B operand only) 256
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    // reduction range: the whole K (fwd / dgrad) or this split's batch rows (wgrad)
    int r_begin = 0, r_end = p.R;
    if (MODE == MODE_WGRAD) { r_begin = min((int)blockIdx.z * p.rows_per_split, p.R); r_end = min(r_begin + p.rows_per_split, p.R); }
    const int n_slabs = (r_end - r_begin + SLAB - 1) / SLAB;
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    constexpr bool A_MN = MODE == MODE_WGRAD, B_MN = MODE != MODE_FWD;
    const int lgb = B_MN ? ilog2_pow2(BN / 8) : 3;
    // tile geometry of slab `sl` (out-of-range elements are zeros: they add nothing).  K-major tiles: the reduction runs along the
    // columns, whose range [rr, rr + 64) is clipped by s.cols = R; MN-major tiles: along the rows, clipped by r_end.
    auto geo_a = [&](int sl) {
        const int rr = r_begin + sl * SLAB;
        return A_MN ? TileGeo{rr, m0, 4, SLAB * (BM / 8), r_end}                      // [64 reduction rows x 128 cols m]
                    : TileGeo{m0, rr, 3, BM * (SLAB / 8), p.a.rows};                  // [128 rows m x 64 reduction cols]
    };
    auto geo_b = [&](int sl) {
        const int rr = r_begin + sl * SLAB;
        return B_MN ? TileGeo{rr, n0, lgb, SLAB * (BN / 8), min(r_end, p.b.rows)}     // [64 reduction rows x BN cols n]
                    : TileGeo{n0, rr, 3, BN * (SLAB / 8), p.b.rows};                  // [BN rows n x 64 reduction cols]
    };
    // one slab of both operands in registers: A 4 chunks of 8 fp32, B 8 chunks of 8 bf16 (BN <= 256) or 4 chunks of 8 fp32 (BN <= 128)
    float4 a_lo[4], a_hi[4];
    float4 b_lo[BB16 ? 1 : 4], b_hi[BB16 ? 1 : 4];
    uint4 b_u[BB16 ? 8 : 1];
    if (n_slabs > 0) {
        const TileGeo ga = geo_a(0), gb = geo_b(0);
        load_f32<4>(a_lo, a_hi, p.a, ga, tid);
        if constexpr (BB16) load_b16<8>(b_u, p.b, gb, tid); else load_f32<4>(b_lo, b_hi, p.b, gb, tid);
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    float bsum = 0.f;        // wgrad: bias gradient of output feature m0 + tid (threads < 128 of the CTAs with blockIdx.y == 0)
    for (int sl = 0; sl < n_slabs; ++sl) {
        const int st = sl & 1, use = sl >> 1;
        uint8_t* sA = smem + st * STAGE_BYTES; uint8_t* sB = sA + A_BYTES;
        if (use > 0) mbar_wait(&bars[st], (uint32_t)(use - 1) & 1u);       // the MMAs that read this stage two slabs ago are done
        {
            const TileGeo ga = geo_a(sl), gb = geo_b(sl);
            store_f32<4>(sA, a_lo, a_hi, p.a, ga, tid);
            if constexpr (BB16) store_b16<8>(sB, b_u, gb, tid); else store_f32<4>(sB, b_lo, b_hi, p.b, gb, tid);
        }
        if (sl + 1 < n_slabs) {             // prefetch: in flight across the barrier, the MMA issue and the next stage wait
            const TileGeo ga = geo_a(sl + 1), gb = geo_b(sl + 1);
            load_f32<4>(a_lo, a_hi, p.a, ga, tid);
            if constexpr (BB16) load_b16<8>(b_u, p.b, gb, tid); else load_f32<4>(b_lo, b_hi, p.b, gb, tid);
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            const uint32_t a_rs = A_MN ? (uint32_t)(BM / 8) * 128u : (uint32_t)(SLAB / 8) * 128u;
            const uint32_t b_rs = B_MN ? (uint32_t)(BN / 8) * 128u : (uint32_t)(SLAB / 8) * 128u;
#pragma unroll
            for (int k = 0; k < SLAB / 16; ++k) {
                const uint64_t ad = A_MN ? make_smem_desc(smem_u32(sA) + k * 2 * a_rs, a_rs, 128u) : make_smem_desc(smem_u32(sA) + k * 2 * 128u, 128u, a_rs);
                const uint64_t bd = B_MN ? make_smem_desc(smem_u32(sB) + k * 2 * b_rs, b_rs, 128u) : make_smem_desc(smem_u32(sB) + k * 2 * 128u, 128u, b_rs);
                umma_bf16(tmem, ad, bd, idesc, (sl > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&bars[st]);
        }
        if (MODE == MODE_WGRAD && p.db && blockIdx.y == 0 && tid < BM) {
            // column sum of the (bf16) dY slab just staged: element (r, c = tid) of the [64 x 128] MN-major tile
            const uint8_t* col = sA + (uint32_t)(tid >> 3) * 128u + (uint32_t)(tid & 7) * 2u;
            float s = 0.f;
#pragma unroll 8
            for (int r = 0; r < SLAB; ++r)
                s += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(col + (uint32_t)(r & 7) * 16u + (uint32_t)(r >> 3) * (uint32_t)(BM / 8) * 128u));
            bsum += s;
        }
    }
    if (n_slabs > 0) {
        const int last = n_slabs - 1;
        mbar_wait(&bars[last & 1], (uint32_t)(last >> 1) & 1u);          // commits are ordered: the last one covers every MMA
        fence_after_sync();
    }
    __syncthreads();          // every MMA has read its operands and the bias sums are done: the stages become the epilogue's transpose scratch
    // ---- epilogue: TMEM lane = output row; each warp owns rows [q * 32, +32) x columns [h * BN/2, +BN/2), in chunks of 32 columns that
    //      it transposes through a private [32][33] shared-memory patch so that every global access is one 128-byte row segment ----
    float* sT = reinterpret_cast<float*>(smem) + warp * (32 * 33);
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int gm0 = m0 + q * 32;
    const int n_rows = min(32, p.Mo - gm0);
    for (int c0 = h * (BN / 2); c0 < (h + 1) * (BN / 2); c0 += 32) {
        const int gn = n0 + c0 + lane;              // this lane's output column for the whole chunk
        const bool col_ok = gn < p.No;
        const int64_t o0 = (int64_t)gm0 * p.No + gn;
        // per-element inputs of the chunk (the accumulate operand / the activation-derivative input): all 32 rows in flight before the
        // TMEM load and the transpose, instead of one dependent global round trip per row
        float pre[32];
        const float* psrc = (MODE == MODE_FWD) ? (p.accumulate ? p.Y : nullptr) : (MODE == MODE_DGRAD ? p.A_prev : nullptr);
        if (psrc) {
#pragma unroll
            for (int k = 0; k < 32; ++k) pre[k] = (col_ok && k < n_rows) ? __ldcg(psrc + o0 + (int64_t)k * p.No) : 0.f;
        }
        if (n_slabs > 0) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(tmem + lane_base + (uint32_t)c0)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) sT[lane * 33 + j] = __uint_as_float(r[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) sT[lane * 33 + j] = 0.f;
        }
        __syncwarp();
        if (col_ok) {
            if (MODE == MODE_FWD) {
                const float bias = p.bias ? __ldg(p.bias + gn) : 0.f;
                float* out = p.Y + o0;
                // the activation switch is hoisted out of the row loop (one uniform branch per chunk, not per element)
                switch (p.act) {
                    case B200RL_ACT_ELU:  fwd_rows<B200RL_ACT_ELU>(out, sT, pre, psrc != nullptr, bias, lane, n_rows, p.No); break;
                    case B200RL_ACT_RELU: fwd_rows<B200RL_ACT_RELU>(out, sT, pre, psrc != nullptr, bias, lane, n_rows, p.No); break;
                    case B200RL_ACT_TANH: fwd_rows<B200RL_ACT_TANH>(out, sT, pre, psrc != nullptr, bias, lane, n_rows, p.No); break;
                    default:              fwd_rows<B200RL_ACT_NONE>(out, sT, pre, psrc != nullptr, bias, lane, n_rows, p.No); break;
                }
            } else if (MODE == MODE_DGRAD) {
                float* out = p.dX + o0;
                switch (psrc ? p.act_prev : B200RL_ACT_NONE) {
                    case B200RL_ACT_ELU:  dgrad_rows<B200RL_ACT_ELU>(out, sT, pre, lane, n_rows, p.No); break;
                    case B200RL_ACT_RELU: dgrad_rows<B200RL_ACT_RELU>(out, sT, pre, lane, n_rows, p.No); break;
                    case B200RL_ACT_TANH: dgrad_rows<B200RL_ACT_TANH>(out, sT, pre, lane, n_rows, p.No); break;
                    default:              dgrad_rows<B200RL_ACT_NONE>(out, sT, pre, lane, n_rows, p.No); break;
                }
            } else {
                float* out = p.dW + (int64_t)blockIdx.z * p.split_stride + o0;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (k < n_rows) out[(int64_t)k * p.No] = sT[k * 33 + lane];
            }
        }
        __syncwarp();       // the patch is rewritten by the next chunk
    }
    if (MODE == MODE_WGRAD && p.db && blockIdx.y == 0 && tid < BM && m0 + tid < p.Mo)
        p.db[(int64_t)blockIdx.z * p.split_stride + m0 + tid] = bsum;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// Output columns per CTA: the widest tile the B operand's register staging allows (256 with a bf16 twin, 128 for fp32 sources), halved
// while the grid would leave SMs idle (fewer than one CTA per SM) -- narrower tiles re-read the A operand from L2, idle SMs cost more.
// a source whose chunk covers all its rows is dense: no per-row division on the device
static inline int dense_or(int rows_per_chunk, int rows) { return rows_per_chunk >= rows ? 0 : rows_per_chunk; }

static inline int pick_bn(int n_out, int m_tiles, int grid_z, int bn_cap) {
    int bn = n_out > 128 ? 256 : (n_out > 64 ? 128 : 64);
    if (bn > bn_cap) bn = bn_cap;
    while (bn > 64 && (int64_t)m_tiles * ((n_out + bn - 1) / bn) * grid_z < 148) bn >>= 1;
    return bn;
}

template <int MODE, bool BB16>
static int launch_gemm_t(const GemmArgs& a, int grid_z, void* stream) {
    constexpr size_t smem = 2 * STAGE_BYTES;
    static bool raised = false;
    if (!raised) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<MODE, BB16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        raised = true;
    }
    dim3 grid((a.Mo + BM - 1) / BM, (a.No + a.BN - 1) / a.BN, grid_z);
    gemm_tc_kernel<MODE, BB16><<<grid, GT, smem, as_stream(stream)>>>(a);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}
template <int MODE>
static int launch_gemm(GemmArgs& a, int grid_z, void* stream) {
    const bool bb16 = a.b.pb != nullptr;
    a.BN = pick_bn(a.No, (a.Mo + BM - 1) / BM, grid_z, bb16 ? 256 : 128);
    return bb16 ? launch_gemm_t<MODE, true>(a, grid_z, stream) : launch_gemm_t<MODE, false>(a, grid_z, stream);
}

}  // namespace

B200RL_EXPORT int b200rl_linear_fwd_tc(const float* X, int rows_per_chunk, int64_t chunk_stride, int64_t x_ld,
                                       const float* norm_mean, const float* norm_std,
                                       const float* W, const void* W_bf16, const float* b, float* Y, int M, int K, int Nout, int act,
                                       int accumulate, void* stream) {
    if (!X || !W || !Y || M <= 0 || K <= 0 || Nout <= 0 || rows_per_chunk <= 0) return B200RL_EINVAL;
    if ((norm_mean == nullptr) != (norm_std == nullptr)) return B200RL_EINVAL;
    GemmArgs a{};
    a.a = Src{X, x_ld, dense_or(rows_per_chunk, M), chunk_stride, norm_mean, norm_std, M, K, nullptr};
    a.b = Src{W, (int64_t)K, 0, 0, nullptr, nullptr, Nout, K, (const __nv_bfloat16*)W_bf16};
    a.Mo = M; a.No = Nout; a.R = K;
    a.bias = b; a.Y = Y; a.act = act; a.accumulate = accumulate;
    return launch_gemm<MODE_FWD>(a, 1, stream);
}

B200RL_EXPORT int b200rl_linear_bwd_data_tc(const float* dY, const float* W, const void* W_bf16, const float* A_prev, float* dX,
                                            int M, int K, int Nout, int act_prev, void* stream) {
    if (!dY || !W || !dX || M <= 0 || K <= 0 || Nout <= 0) return B200RL_EINVAL;
    GemmArgs a{};
    a.a = Src{dY, (int64_t)Nout, 0, 0, nullptr, nullptr, M, Nout, nullptr};            // [m, n]: reduction index n contiguous -> K-major
    a.b = Src{W, (int64_t)K, 0, 0, nullptr, nullptr, Nout, K, (const __nv_bfloat16*)W_bf16};   // [n, k]: rows = reduction n, cols = outputs k -> MN-major
    a.Mo = M; a.No = K; a.R = Nout;
    a.A_prev = A_prev; a.dX = dX; a.act_prev = act_prev;
    return launch_gemm<MODE_DGRAD>(a, 1, stream);
}

B200RL_EXPORT int b200rl_linear_bwd_weight_tc(const float* dY, const float* X, int rows_per_chunk, int64_t chunk_stride,
                                              int64_t x_ld, const float* norm_mean, const float* norm_std,
                                              float* dW_part, float* db_part, int64_t split_stride, int M, int K, int Nout,
                                              int n_splits, void* stream) {
    if (!dY || !X || !dW_part || M <= 0 || K <= 0 || Nout <= 0 || n_splits <= 0 || rows_per_chunk <= 0) return B200RL_EINVAL;
    if ((norm_mean == nullptr) != (norm_std == nullptr)) return B200RL_EINVAL;
    GemmArgs a{};
    a.a = Src{dY, (int64_t)Nout, 0, 0, nullptr, nullptr, M, Nout, nullptr};            // [r, n]: rows = reduction r, cols = outputs n -> MN-major
    a.b = Src{X, x_ld, dense_or(rows_per_chunk, M), chunk_stride, norm_mean, norm_std, M, K, nullptr};   // [r, k]: rows = reduction r, cols = outputs k -> MN-major
    a.Mo = Nout; a.No = K; a.R = M;
    a.dW = dW_part; a.db = db_part; a.split_stride = split_stride;
    int rps = (M + n_splits - 1) / n_splits;
    a.rows_per_split = ((rps + SLAB - 1) / SLAB) * SLAB;       // trailing splits may be empty: they write zeros, which the reducer expects
    return launch_gemm<MODE_WGRAD>(a, n_splits, stream);
}


// fp32 -> bf16 copy of the flat parameter arena (same offsets): the weight operand of the layer-wise GEMMs, refreshed once per optimiser step
namespace {
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = __float2bfloat16_rn(__ldg(src + i));
}
}  // namespace
B200RL_EXPORT int b200rl_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream) {
    if (!src || !dst_bf16 || n <= 0) return B200RL_EINVAL;
    const int blocks = (int)((n + 1023) / 1024 < 1184 ? (n + 1023) / 1024 : 1184);
    cast_bf16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, (__nv_bfloat16*)dst_bf16, n);
    B200RL_LAUNCH_CHECK();
    return B200RL_OK;
}
