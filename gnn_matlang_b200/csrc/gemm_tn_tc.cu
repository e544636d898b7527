// Weight-gradient contraction on tcgen05:  C[Ka, Nb] = A[M, Ka]^T * B[M, Nb]   (Ka <= 32 input features, Nb <= 256
// = K supports x 32 columns of the transposed aggregate; M = nodes of the batch, 10^5..10^6).
//
// The contraction runs over the ROWS of both row-major operands, i.e. both tensor-core operands are MN-major.  For 32-bit
// elements the only MN-major shared-memory layout tcgen05 accepts is SWIZZLE_128B_BASE32B (layout type 1): atoms of
// 4 k-rows x 128 bytes (32 columns), the 32-byte chunk index XOR-ed with k % 4; stride-byte-offset = distance between
// k-atoms (512 B), leading-byte-offset = distance between 32-column atoms (scratch/umma_mn_probe.cu pins this on the
// B200: every other layout/type returns zeros).
//
// 3xTF32 with two instructions per 8 rows:  A' = [x_hi ; x_lo] on the M side (M = 64), B' = G_hi resp. G_lo (N = Nb):
//     D1 = [x_hi; x_lo]^T G_hi   (rows 0-31: hi*hi, rows 32-63: lo*hi)      D2 = [x_hi; x_lo]^T G_lo  (rows 0-31: hi*lo)
//     C  = D1[0:32] + D1[32:64] + D2[0:32]
// (the tensor core truncates the raw FP32 bits to TF32 = "hi"; lo = v - trunc(v) is exact in FP32).  One persistent CTA per
// SM owns a contiguous row range; one thread streams 32-row stages with TMA (SWIZZLE_128B_ATOM_32B boxes = the raw/hi
// planes, three stages ahead), 8 warps compute the lo planes from them, one thread issues the MMAs, two
// alternating accumulator sets keep every TMEM accumulation chain at <= M / (148 * 16) adds.  Per-CTA partials go to the
// workspace and k_reduce_partials (gemm.cu) adds them in a fixed order -- deterministic, no atomics.
#include "tc_common.cuh"

namespace gnnml3 {

constexpr int TT_ROWS = 32, TT_HI = 3, TT_LO = 3;             // raw / lo slots
constexpr int TT_PRODUCERS = 256, TT_THREADS = TT_PRODUCERS + 64;   // 8 split warps + MMA warp + TMA warp
constexpr int TT_PLANE = TT_ROWS * 128;                       // one 32-column atom column of a stage: 4 KB
constexpr int TT_NPL = 8;                                     // B planes per stage (Nb <= 256)
constexpr int TT_HALF = (1 + TT_NPL) * TT_PLANE;              // x plane + 8 G planes of one precision half: 36 KB
constexpr size_t TT_SMEM = (size_t)(TT_HI + TT_LO) * TT_HALF + 1024;   // hi ring + lo ring = 216 KB
constexpr int TT_OUT_LD = 260;                                // epilogue staging pitch (floats)

__device__ __forceinline__ uint64_t make_mnmajor_b32_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;       // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;       // SWIZZLE_128B_BASE32B
    return d;
}
// byte offset of the 16-byte chunk holding columns m..m+3 (m % 4 == 0) of stage row k, inside an operand's plane group
__device__ __forceinline__ uint32_t tt_off(int k, int m) {
    return (uint32_t)((m >> 5) * TT_PLANE + (k >> 2) * 512 + (k & 3) * 128 + (((((m & 31) >> 3) ^ (k & 3))) << 5) + (m & 7) * 4);
}
__device__ __forceinline__ float tt_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__global__ void __launch_bounds__(TT_THREADS, 1)
k_gemm_tn_tc(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ P,
             int64_t M, int Ka, int Nb, int64_t rows_per_cta) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    __shared__ uint64_t raw[TT_HI], full[TT_HI], mdone[TT_HI], done;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
    const int64_t rbeg = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t rend = min(M, rbeg + rows_per_cta);
    const int nst = rbeg < rend ? (int)((rend - rbeg + TT_ROWS - 1) / TT_ROWS) : 0;
    const int Npad = (Nb + 31) & ~31;
    const int nch = Npad >> 2;                                 // 16-byte chunks per B row

    if (t == 0) {
        for (int s = 0; s < TT_HI; ++s) { mbar_init(&raw[s], 1); mbar_init(&full[s], TT_PRODUCERS / 32); mbar_init(&mdone[s], 1); }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TT_PRODUCERS / 32) tmem_alloc(&tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tslot;

    uint8_t* hi_ring = smem;                                   // TT_HI x [x_hi | G_hi x 8]
    uint8_t* lo_ring = smem + (size_t)TT_HI * TT_HALF;         // TT_LO x [x_lo | G_lo x 8]
    if (warp < TT_PRODUCERS / 32) {
        // ------------------------------------------------------------------ split warps: lo = v - trunc(v), same offsets
        const int xk = t >> 3, xc = (t & 7) * 4;
        const uint32_t xoff = tt_off(xk, xc);
        bool gon[8];
        uint32_t goff[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = t + TT_PRODUCERS * u;
            const int k = i / nch, c = (i - k * nch) * 4;
            gon[u] = k < TT_ROWS;
            goff[u] = TT_PLANE + tt_off(k & (TT_ROWS - 1), c);
        }
        for (int it = 0; it < nst; ++it) {
            const int s = it % TT_HI;
            mbar_wait(&raw[s], (it / TT_HI) & 1);                                                    // TMA boxes of stage it landed
            if (it >= TT_LO) mbar_wait(&mdone[(it - TT_LO) % TT_HI], ((it - TT_LO) / TT_HI) & 1);   // lo slot is free again
            const uint8_t* hs = hi_ring + (size_t)s * TT_HALF;
            uint8_t* ls = lo_ring + (size_t)(it % TT_LO) * TT_HALF;
            {
                const float4 v = *reinterpret_cast<const float4*>(hs + xoff);
                *reinterpret_cast<float4*>(ls + xoff) = make_float4(tt_lo(v.x), tt_lo(v.y), tt_lo(v.z), tt_lo(v.w));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (gon[u]) {
                    const float4 v = *reinterpret_cast<const float4*>(hs + goff[u]);
                    *reinterpret_cast<float4*>(ls + goff[u]) = make_float4(tt_lo(v.x), tt_lo(v.y), tt_lo(v.z), tt_lo(v.w));
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);             // one arrival per warp (arrivals on one barrier serialise)
        }
    } else if (warp == TT_PRODUCERS / 32 + 1) {
        // ------------------------------------------------------------------ TMA issuer: x box + Npad / 32 G boxes per stage
        if (lane == 0) {
            const int npl = Npad >> 5;
            const uint32_t bytes = (uint32_t)(1 + npl) * TT_PLANE;
            for (int it = 0; it < nst; ++it) {
                const int s = it % TT_HI;
                if (it >= TT_HI) mbar_wait(&mdone[s], ((it / TT_HI) - 1) & 1);      // MMAs that read this raw slot are done
                uint8_t* st = hi_ring + (size_t)s * TT_HALF;
                const int r0 = (int)(rbeg + (int64_t)it * TT_ROWS);
                mbar_arrive_expect_tx(&raw[s], bytes);
                tma_load_2d(st, &mapA, &raw[s], 0, r0);
                for (int pl = 0; pl < npl; ++pl) tma_load_2d(st + (size_t)(1 + pl) * TT_PLANE, &mapB, &raw[s], 32 * pl, r0);
            }
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------------ MMA issuer (warp 8)
        const uint32_t idesc = make_idesc_tf32_mn(64, Npad, true, true);
        for (int it = 0; it < nst; ++it) {
            const int s = it % TT_HI;
            mbar_wait(&full[s], (it / TT_HI) & 1);
            tc_fence_after();
            const uint32_t hb = smem_u32(hi_ring + (size_t)s * TT_HALF);
            const uint32_t lo_delta = smem_u32(lo_ring + (size_t)(it % TT_LO) * TT_HALF) - hb;   // x_lo plane - x_hi plane
            const uint32_t d1 = tm + (uint32_t)(it & 1) * 256u, d2 = d1 + (16u << 16);
#pragma unroll
            for (int j = 0; j < TT_ROWS / 8; ++j) {
                const uint64_t da = make_mnmajor_b32_desc(hb + j * 1024, lo_delta, 512);             // M atoms: x_hi, x_lo
                const uint64_t dbh = make_mnmajor_b32_desc(hb + TT_PLANE + j * 1024, TT_PLANE, 512);
                const uint64_t dbl = make_mnmajor_b32_desc(hb + lo_delta + TT_PLANE + j * 1024, TT_PLANE, 512);
                const uint32_t acc = (it >= 2 || j > 0) ? 1u : 0u;
                umma_tf32(d1, da, dbh, idesc, acc);
                umma_tf32(d2, da, dbl, idesc, acc);
            }
            umma_commit(&mdone[s]);
        }
        if (nst > 0) umma_commit(&done);
    }

    // ---------------------------------------------------------------------- epilogue: TMEM -> smem (3 terms) -> partial
    float* outs = reinterpret_cast<float*>(smem);             // [3][32][TT_OUT_LD], reuses the stage ring
    if (nst > 0) {
        mbar_wait(&done, 0);
        tc_fence_after();
        // every MMA that read the stage ring has retired (commit -> `done`) and every split warp finished its reads before its last
        // arrival, so `outs` may reuse the ring; the CTA barrier states that ordering in a form compute-sanitizer's racecheck can see
        // (it does not follow mbarrier / tcgen05.commit edges and reported the reuse as a hazard)
        __syncthreads();
        if (warp < 4) {
            // lanes 0-15 of warp w: D1 rows 16w..16w+15; lanes 16-31: D2 rows 16w..16w+15 (accumulator at lane offset 16)
            const int mrow = 16 * warp + (lane & 15);
            const int term = (warp < 2) ? (lane < 16 ? 0 : 1) : (lane < 16 ? 2 : 3);
            float* dst = outs + ((size_t)term * 32 + (mrow & 31)) * TT_OUT_LD;
            for (int c0 = 0; c0 < Npad; c0 += 32) {
                float v[32];
                tmem_ld32(tm + ((uint32_t)(32 * warp) << 16) + c0, v);
                if (nst > 1) {
                    float w[32];
                    tmem_ld32(tm + ((uint32_t)(32 * warp) << 16) + 256u + c0, w);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += w[i];
                }
                if (term < 3) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    float* Pz = P + (size_t)blockIdx.x * Ka * Nb;
    for (int i = t; i < Ka * Nb; i += TT_THREADS) {
        const int a = i / Nb, c = i - a * Nb;
        float v = 0.f;
        if (nst > 0)
            v = (outs[(size_t)a * TT_OUT_LD + c] + outs[((size_t)64 + a) * TT_OUT_LD + c]) + outs[((size_t)32 + a) * TT_OUT_LD + c];
        Pz[i] = v;
    }
    if (warp == TT_PRODUCERS / 32) tmem_dealloc(tm, 512);
}

// shape-only eligibility (the workspace query has no pointers); pointer alignment is checked at launch
bool gemm_tn_tc_shape_ok(int64_t M, int Ka, int Nb) { return Ka >= 1 && Ka <= 32 && Nb > 8 && M >= 8192; }   // Nb > 256: 256-column launches
int gemm_tn_tc_parts(int64_t M) {
    const int64_t tiles = (M + TT_ROWS - 1) / TT_ROWS;
    return (int)(tiles < kNumSMs ? tiles : kNumSMs);
}
bool gemm_tn_tc_ok(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int Ka, int Nb) {
    return gemm_tn_tc_shape_ok(M, Ka, Nb) && lda % 4 == 0 && ldb % 4 == 0 && (uintptr_t)A % 16 == 0 && (uintptr_t)B % 16 == 0;
}
// row-major FP32 [rows, cols] (row stride ld) -> boxes of 32 columns x TT_ROWS rows in the 128B-swizzle / 32B-atom pattern
// (= UMMA SWIZZLE_128B_BASE32B); columns / rows outside the matrix are zero-filled by the TMA unit
static int tt_make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
    PFN_encodeTiled enc = get_encoder();
    if (!enc) return set_err(GNNML3_ERR_CUDA, "gemm_tn: cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)TT_ROWS};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(GNNML3_ERR_CUDA, "gemm_tn: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return GNNML3_OK;
}

// P: [parts][Ka * Nb] partials
int gemm_tn_tc_launch(const float* A, int64_t lda, const float* B, int64_t ldb, float* P, int64_t M, int Ka, int Nb,
                      cudaStream_t st) {
    static bool configured[64] = {};
    if (auto once_ = first_use_on_device(configured))
        GNNML3_CUDA(cudaFuncSetAttribute(k_gemm_tn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TT_SMEM));
    CUtensorMap mapA, mapB;
    int rc;
    if ((rc = tt_make_map(&mapA, A, M, Ka, lda))) return rc;
    if ((rc = tt_make_map(&mapB, B, M, Nb, ldb))) return rc;
    const int parts = gemm_tn_tc_parts(M);
    int64_t rpc = (M + parts - 1) / parts;
    rpc = (rpc + TT_ROWS - 1) / TT_ROWS * TT_ROWS;
    k_gemm_tn_tc<<<parts, TT_THREADS, TT_SMEM, st>>>(mapA, mapB, P, M, Ka, Nb, rpc);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

}  // namespace gnnml3
