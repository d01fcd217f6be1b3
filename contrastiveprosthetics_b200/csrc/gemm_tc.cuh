// Tensor-core engine: fp32-accurate GEMMs on tcgen05 (kind::tf32) with the 3xTF32 split
//   x = hi + lo, hi = tf32(x), lo = tf32(x - hi);   a.b ~= hi_a.hi_b + hi_a.lo_b + lo_a.hi_b
// (dropped term lo.lo <= 2^-22 |a||b|), accumulated in fp32 in TMEM.  Operands arrive pre-split as
// separate fp32 planes (written by the producing elementwise kernels), so the main loop is pure
// TMA -> shared memory -> tcgen05.mma with no register traffic.
//
//   gemm_tc_nt : C[M,N] = act(A[M,K] . B[N,K]^T + bias) + per-column sum / sum-of-squares partials
//                (both operands K-major, 128B-swizzled tiles).           forward + data gradient
//   gemm_tc_tn : P[z][Mo,No] = sum_{r in slab z} G[r,Mo] * A[r,No]       weight gradient (split-K)
//                (both operands MN-major, 128B-swizzled tiles)
//
// CTA = 6 warps: warp 0 TMA producer (one elected lane), warp 1 MMA issuer (one elected lane; owns
// the TMEM allocation), warps 2-5 epilogue (TMEM -> registers -> global).  Persistent: grid =
// min(#tiles, 148); two TMEM accumulators (2 x 128 columns) so the epilogue of tile i overlaps the
// main loop of tile i+1; 3-stage shared-memory ring (A_hi, A_lo, B_hi, B_lo: 4 x 16 KB per stage).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace tcg {

constexpr int BM = 128, BN = 128, BK = 32;        // BK floats = 128 bytes = one swizzle row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;           // 16 KB per operand plane per stage
constexpr int STAGE_BYTES = 4 * TILE_BYTES;       // A_hi, A_lo, B_hi, B_lo
constexpr int UMMA_K = 8;                         // tf32: 32 bytes of K per instruction
constexpr int THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr int TMEM_COLS = 512;                    // 2 buffers x (main + correction accumulator) x 128 fp32 columns
constexpr int ACC_COLS = 2 * BN;                  // columns per buffer

struct Smem {
    uint64_t full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    float csum[4][BN];
    float csq[4][BN];
};
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + (int)sizeof(Smem);

struct NtArgs {
    float* C; int ldc;
    const float* bias;
    float* psum; float* psq;        // [ceil(M/128)][N] or null
    int64_t M; int N; int K;
    int relu;
};

// warp-transposing reduction: on return v[0] of lane l = sum over the 32 lanes of their v[l]
__device__ __forceinline__ void warp_col_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_nt_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                  const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                  const NtArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Smem* sm = reinterpret_cast<Smem*>(tiles + STAGES * STAGE_BYTES);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int tiles_n = g.N / BN;
    const int64_t tiles_m = (g.M + BM - 1) / BM;
    const int64_t n_tiles = tiles_m * tiles_n;
    const int kblocks = g.K / BK;

    if (warp == 0 && tc::elect_one()) {
        tc::prefetch_tmap(&tm_a_hi); tc::prefetch_tmap(&tm_a_lo);
        tc::prefetch_tmap(&tm_b_hi); tc::prefetch_tmap(&tm_b_lo);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&sm->full[s], 1); tc::mbar_init(&sm->empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&sm->tmem_full[a], 1); tc::mbar_init(&sm->tmem_empty[a], 4); }
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(&sm->tmem_base, TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (tc::elect_one()) {
            int s = 0; uint32_t ph = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int m0 = (int)(t / tiles_n) * BM, n0 = (int)(t % tiles_n) * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    tc::mbar_wait(&sm->empty[s], ph ^ 1);
                    uint8_t* st = tiles + s * STAGE_BYTES;
                    tc::mbar_expect_tx(&sm->full[s], STAGE_BYTES);
                    tc::tma_load_2d(st + 0 * TILE_BYTES, &tm_a_hi, &sm->full[s], kb * BK, m0);
                    tc::tma_load_2d(st + 1 * TILE_BYTES, &tm_a_lo, &sm->full[s], kb * BK, m0);
                    tc::tma_load_2d(st + 2 * TILE_BYTES, &tm_b_hi, &sm->full[s], kb * BK, n0);
                    tc::tma_load_2d(st + 3 * TILE_BYTES, &tm_b_lo, &sm->full[s], kb * BK, n0);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::idesc_tf32(BM, BN, 0, 0);
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                tc::mbar_wait(&sm->tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                // two accumulators per tile: the hi.hi chain and the (small) correction chain are kept
                // apart so that the accumulator's round-toward-zero steps on the big chain are 1/3 as many
                const uint32_t d = tmem_base + acc * ACC_COLS;
                const uint32_t dc = d + BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    tc::mbar_wait(&sm->full[s], ph);
                    tc::tc_fence_after();
                    const uint32_t base = tc::smem_u32(tiles + s * STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t ko = k * UMMA_K * 4;                       // bytes along K inside the swizzle row
                        const uint64_t a_hi = tc::smem_desc_sw128(base + 0 * TILE_BYTES + ko, 16, 1024);
                        const uint64_t a_lo = tc::smem_desc_sw128(base + 1 * TILE_BYTES + ko, 16, 1024);
                        const uint64_t b_hi = tc::smem_desc_sw128(base + 2 * TILE_BYTES + ko, 16, 1024);
                        const uint64_t b_lo = tc::smem_desc_sw128(base + 3 * TILE_BYTES + ko, 16, 1024);
                        tc::mma_tf32(dc, a_lo, b_hi, idesc, (kb | k) != 0);
                        tc::mma_tf32(dc, a_hi, b_lo, idesc, 1);
                        tc::mma_tf32(d, a_hi, b_hi, idesc, (kb | k) != 0);
                    }
                    tc::mma_commit(&sm->empty[s]);                                // frees the stage when the MMAs retire
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                tc::mma_commit(&sm->tmem_full[acc]);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp % 4;                      // TMEM lane quadrant this warp may access
        const int et = threadIdx.x - 64;             // 0..127
        int it = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const int64_t tile_m = t / tiles_n;
            const int n0 = (int)(t % tiles_n) * BN;
            const int64_t row = tile_m * BM + q * 32 + lane;
            const bool row_ok = row < g.M;
            tc::mbar_wait(&sm->tmem_full[acc], (it >> 1) & 1);
            tc::tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                float v[32], vc[32];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + c * 32;
                tc::tmem_ld32(ta, v);
                tc::tmem_ld32(ta + BN, vc);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += vc[j];
                const int col = n0 + c * 32;
                if (g.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col + j));
                        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                    }
                }
                if (g.relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (row_ok) {
                    float4* dst = reinterpret_cast<float4*>(g.C + row * g.ldc + col);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                if (g.psum) {
                    float sq[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        v[j] = row_ok ? v[j] : 0.f;
                        sq[j] = v[j] * v[j];
                    }
                    warp_col_reduce32(v, lane);
                    warp_col_reduce32(sq, lane);
                    sm->csum[q][c * 32 + lane] = v[0];
                    sm->csq[q][c * 32 + lane] = sq[0];
                }
            }
            // accumulator drained -> hand it back to the MMA warp
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sm->tmem_empty[acc]);
            if (g.psum) {
                tc::named_bar_sync(1, EPI_THREADS);
                const float s = sm->csum[0][et] + sm->csum[1][et] + sm->csum[2][et] + sm->csum[3][et];
                const float qq = sm->csq[0][et] + sm->csq[1][et] + sm->csq[2][et] + sm->csq[3][et];
                g.psum[tile_m * g.N + n0 + et] = s;
                g.psq[tile_m * g.N + n0 + et] = qq;
                tc::named_bar_sync(1, EPI_THREADS);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------- weight gradient
// P[z][o, c] = sum over rows r of slab z of G[r, o] * A[r, c].  Both operands are MN-major: a stage
// holds, per plane, 4 blocks of [32 rows r][32 floats] (one 3-D TMA box {32, 32, 4}); MMA K-step j
// reads the 8 rows at +j*1024 B (two 4-row swizzle groups, SBO = 512 B), the four 32-wide MN blocks
// are LBO = 4096 B apart.  MN-major tf32 operands must use the 32-byte-atom 128B swizzle.
// The fp32 accumulators in TMEM are rounded toward zero at every MMA, so a chain over ~20k rows
// would carry a ~1e-4 bias: the chain is cut every CHUNK_KB k-blocks (512 rows) and the epilogue
// threads keep the running sum in registers (round-to-nearest adds), 128 per thread.
constexpr int CHUNK_KB = 16;

struct TnArgs {
    float* P;                 // [splits][Mo][No]
    int Mo, No;
    int64_t R;                // total rows
    int64_t rows_per_split;   // multiple of BK
};

__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_tn_kernel(const __grid_constant__ CUtensorMap tm_g_hi, const __grid_constant__ CUtensorMap tm_g_lo,
                  const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                  const TnArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Smem* sm = reinterpret_cast<Smem*>(tiles + STAGES * STAGE_BYTES);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int o0 = blockIdx.y * BM, c0 = blockIdx.x * BN;
    const int64_t r_begin = (int64_t)blockIdx.z * g.rows_per_split;
    const int64_t r_end = min(g.R, r_begin + g.rows_per_split);
    const int kblocks = (int)((r_end - r_begin + BK - 1) / BK);
    const int chunks = (kblocks + CHUNK_KB - 1) / CHUNK_KB;

    if (warp == 0 && tc::elect_one()) {
        tc::prefetch_tmap(&tm_g_hi); tc::prefetch_tmap(&tm_g_lo);
        tc::prefetch_tmap(&tm_a_hi); tc::prefetch_tmap(&tm_a_lo);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&sm->full[s], 1); tc::mbar_init(&sm->empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&sm->tmem_full[a], 1); tc::mbar_init(&sm->tmem_empty[a], 4); }
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(&sm->tmem_base, TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;

    if (warp == 0) {
        if (tc::elect_one()) {
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                tc::mbar_wait(&sm->empty[s], ph ^ 1);
                uint8_t* st = tiles + s * STAGE_BYTES;
                const int r = (int)(r_begin + (int64_t)kb * BK);
                tc::mbar_expect_tx(&sm->full[s], STAGE_BYTES);
                tc::tma_load_3d(st + 0 * TILE_BYTES, &tm_g_hi, &sm->full[s], 0, r, o0 / 32);
                tc::tma_load_3d(st + 1 * TILE_BYTES, &tm_g_lo, &sm->full[s], 0, r, o0 / 32);
                tc::tma_load_3d(st + 2 * TILE_BYTES, &tm_a_hi, &sm->full[s], 0, r, c0 / 32);
                tc::tma_load_3d(st + 3 * TILE_BYTES, &tm_a_lo, &sm->full[s], 0, r, c0 / 32);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::idesc_tf32(BM, BN, 1, 1);
            int s = 0; uint32_t ph = 0;
            int kb = 0;
            for (int ch = 0; ch < chunks; ++ch) {
                const int acc = ch & 1;
                tc::mbar_wait(&sm->tmem_empty[acc], ((ch >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t d = tmem_base + acc * ACC_COLS;
                const uint32_t dc = d + BN;
                const int kb_end = min(kblocks, kb + CHUNK_KB);
                for (bool first = true; kb < kb_end; ++kb) {
                    tc::mbar_wait(&sm->full[s], ph);
                    tc::tc_fence_after();
                    const uint32_t base = tc::smem_u32(tiles + s * STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t ko = k * 1024;                 // 8 rows of 128 B
                        const uint64_t g_hi = tc::smem_desc(base + 0 * TILE_BYTES + ko, 4096, 512, 1);
                        const uint64_t g_lo = tc::smem_desc(base + 1 * TILE_BYTES + ko, 4096, 512, 1);
                        const uint64_t a_hi = tc::smem_desc(base + 2 * TILE_BYTES + ko, 4096, 512, 1);
                        const uint64_t a_lo = tc::smem_desc(base + 3 * TILE_BYTES + ko, 4096, 512, 1);
                        const uint32_t accum = (first && k == 0) ? 0u : 1u;
                        tc::mma_tf32(dc, g_lo, a_hi, idesc, accum);
                        tc::mma_tf32(dc, g_hi, a_lo, idesc, 1);
                        tc::mma_tf32(d, g_hi, a_hi, idesc, accum);
                    }
                    first = false;
                    tc::mma_commit(&sm->empty[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                tc::mma_commit(&sm->tmem_full[acc]);
            }
        }
    } else {
        const int q = warp % 4;
        float sum[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) sum[j] = 0.f;
        for (int ch = 0; ch < chunks; ++ch) {
            const int acc = ch & 1;
            tc::mbar_wait(&sm->tmem_full[acc], (ch >> 1) & 1);
            tc::tc_fence_after();
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                float v[32], vc[32];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + c * 32;
                tc::tmem_ld32(ta, v);
                tc::tmem_ld32(ta + BN, vc);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[c * 32 + j] += v[j] + vc[j];
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sm->tmem_empty[acc]);
        }
        float* dst = g.P + ((int64_t)blockIdx.z * g.Mo + o0 + q * 32 + lane) * g.No + c0;
#pragma unroll
        for (int j = 0; j < BN; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major fp32 [rows, cols] (leading dimension ld floats), box {32 cols, box_rows}, 128B swizzle,
// out-of-range elements read as zero
inline int make_tmap_2d(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return CP_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CP_OK : CP_ERR_ARG;
}

// row-major fp32 [rows, cols] viewed as (32 cols, rows, cols/32 blocks): box {32, 32 rows, 4 blocks}
// lands in shared memory as 4 x [32 rows][128 B], i.e. the MN-major 128B-swizzle canonical layout
inline int make_tmap_mn(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return CP_ERR_UNSUPPORTED;
    cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
    cuuint32_t box[3] = {32, 32, 4};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CP_OK : CP_ERR_ARG;
}

// returns the number of splits written to P ([splits][Mo][No]) through *splits_out
inline int launch_tn(const float* G_hi, const float* G_lo, int ldg, int Mo, const float* A_hi, const float* A_lo,
                     int lda, int No, int64_t R, float* P, size_t p_capacity_elems, int* splits_out,
                     cudaStream_t st) {
    if (Mo % BM != 0 || No % BN != 0 || ldg % 4 != 0 || lda % 4 != 0 || R <= 0) return CP_ERR_ARG;
    CUtensorMap tg_hi, tg_lo, ta_hi, ta_lo;
    int rc;
    if ((rc = make_tmap_mn(&tg_hi, G_hi, R, Mo, ldg)) != CP_OK) return rc;
    if ((rc = make_tmap_mn(&tg_lo, G_lo, R, Mo, ldg)) != CP_OK) return rc;
    if ((rc = make_tmap_mn(&ta_hi, A_hi, R, No, lda)) != CP_OK) return rc;
    if ((rc = make_tmap_mn(&ta_lo, A_lo, R, No, lda)) != CP_OK) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    const int tiles = (Mo / BM) * (No / BN);
    int S = CP_NUM_SMS / tiles;                                   // one CTA per SM, one wave
    const int64_t max_s = cp_cdiv(R, (int64_t)BK * CHUNK_KB);
    if (S > max_s) S = (int)max_s;
    const int64_t cap = (int64_t)(p_capacity_elems / ((size_t)Mo * No));
    if (S > cap) S = (int)cap;
    if (S < 1) S = 1;
    const int64_t rps = cp_cdiv(cp_cdiv(R, S), BK) * BK;
    S = (int)cp_cdiv(R, rps);
    TnArgs g{P, Mo, No, R, rps};
    gemm_tc_tn_kernel<<<dim3(No / BN, Mo / BM, S), THREADS, SMEM_BYTES, st>>>(tg_hi, tg_lo, ta_hi, ta_lo, g);
    CP_CHECK_LAUNCH();
    *splits_out = S;
    return CP_OK;
}

inline int launch_nt(const float* A_hi, const float* A_lo, int64_t M, int K, int lda, const float* B_hi,
                     const float* B_lo, int N, int ldb, const float* bias, float* C, int ldc, float* psum,
                     float* psq, int relu, cudaStream_t st) {
    if (K % BK != 0 || N % BN != 0 || lda % 4 != 0 || ldb % 4 != 0) return CP_ERR_ARG;
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    int rc;
    if ((rc = make_tmap_2d(&ta_hi, A_hi, M, K, lda, BM)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&ta_lo, A_lo, M, K, lda, BM)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&tb_hi, B_hi, N, K, ldb, BN)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&tb_lo, B_lo, N, K, ldb, BN)) != CP_OK) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    NtArgs g{C, ldc, bias, psum, psq, M, N, K, relu};
    const int64_t n_tiles = cp_cdiv(M, BM) * (N / BN);
    const int grid = (int)(n_tiles < CP_NUM_SMS ? n_tiles : CP_NUM_SMS);
    gemm_tc_nt_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, g);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

}  // namespace tcg

// x -> (hi, lo) planes: hi = rna_tf32(x), lo = rna_tf32(x - hi)
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = tf32_rna(x);
    lo = tf32_rna(x - hi);
}
__device__ __forceinline__ void split_tf32(const float4& x, float4& hi, float4& lo) {
    split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y);
    split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
}

__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t n4) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += (int64_t)gridDim.x * blockDim.x) {
        float4 h, l;
        split_tf32(__ldg(reinterpret_cast<const float4*>(x) + v), h, l);
        reinterpret_cast<float4*>(hi)[v] = h;
        reinterpret_cast<float4*>(lo)[v] = l;
    }
}
