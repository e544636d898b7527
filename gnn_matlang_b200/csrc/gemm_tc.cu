// Blackwell-native tall-skinny GEMM:  C[M, Nc] = A[M, Kc] * B[Kc, Nc] (+ bias) (+ ReLU)  on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in tensor memory), FP32-grade through the 3xTF32 split.
//
// Used for the SpectConv projection sum_k P_k(x) W_k as ONE contraction over K*Fi (reference libs/spect_conv.py:80,
// :93-94) and for the dx / dH contractions of its backward.  M (nodes) is huge, Nc and Kc are small, so the kernel
// streams A from HBM exactly once and everything else stays on chip.
//
// Persistent, warp-specialised CTA (one per SM):
//   warp  0    TMA producer: one lane issues cp.async.bulk.tensor loads of the raw A tile [128 x 32 FP32] and of the
//                          (tiny, host-pre-split, transposed) weight planes Bt_hi / Bt_lo straight into the UMMA
//                          canonical K-major SWIZZLE_128B layout (zero fill outside the matrix), STAGES deep.
//   warps 1-4  splitters : a = hi + lo with hi = the TF32 truncation the tensor core itself applies to the raw plane,
//                          so only lo = a - trunc(a) has to be materialised: one LDS.128 / STS.128 pair per 16 bytes at
//                          the SAME (swizzled) offset -- no index arithmetic, no second copy of hi.
//   warp  5    MMA issuer: one elected lane issues, per 32-wide k-block, 4 x 3 tcgen05.mma (lo*hi + hi*lo + hi*hi)
//                          into a TMEM accumulator; tcgen05.commit releases the smem stage / publishes the chunk.
//   warps 6-9  epilogue  : tcgen05.ld the accumulator chunk (32 lanes x BN columns per warp), fold it into FP32
//                          registers with round-to-nearest adds (the tensor core's own accumulate truncates, which
//                          would bias long contractions), finally bias / ReLU and 128-bit row stores.
// 256 TMEM columns hold 4-8 accumulator buffers, so the MMAs run up to a tile ahead of the drain; a 4-stage mbarrier
// ring (raw landed -> split done -> MMAs retired) decouples TMA, splitters and the tensor core.
#include "tc_common.cuh"

namespace gnnml3 {

constexpr int TC_BM = 128, TC_BK = 32, TC_STAGES = 4;
constexpr int TC_SPLITTERS = 128, TC_EPILOGUE = 128;
constexpr int TC_THREADS = 32 + TC_SPLITTERS + 32 + TC_EPILOGUE;   // 320: TMA | splitters | MMA | epilogue
constexpr int TC_A_BYTES = TC_BM * 128;                       // one [128 x 32] FP32 plane

template <int BN>
struct TCCfg {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = 2 * TC_A_BYTES + 2 * B_BYTES;
    static constexpr int NBUF = 256 / BN;          // TMEM accumulator buffers (8 x 32 or 4 x 64 columns): the MMAs may run
    static constexpr int TMEM_COLS = 256;          // a whole tile ahead while the epilogue is still storing the previous one
    static constexpr uint32_t TX_BYTES = TC_A_BYTES + 2 * B_BYTES;    // bytes TMA delivers per stage
    static constexpr size_t SMEM = (size_t)TC_STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 512 /*barriers*/;
};

// instruction descriptor: D = F32, A = B = TF32, both K-major, N = BN, M = 128
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc_tf32() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------- weight planes
// Bt_hi / Bt_lo [Npad][Kpad]: transposed (K-major), zero padded, pre-split weights.  Tiny (<= 256 x 2560).
__global__ void k_prep_weights_tc(const float* __restrict__ B, int64_t ldb, int Kc, int Nc, int Kpad, int Npad,
                                  float* __restrict__ hi, float* __restrict__ lo) {
    const int total = Npad * Kpad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / Kpad, k = i % Kpad;
        const float v = (n < Nc && k < Kc) ? __ldg(B + (int64_t)k * ldb + n) : 0.f;
        const float h = tf32_rn(v);
        hi[i] = h;
        lo[i] = v - h;
    }
}

// ---------------------------------------------------------------------------------------------- the GEMM
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_nn_tc(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBhi,
             const __grid_constant__ CUtensorMap mapBlo, const float* __restrict__ bias, float* __restrict__ C, int64_t ldc,
             int64_t M, int Nc, int Kc, int epi, int chunk_kb, int n_mtiles, int b_res) {
    using Cfg = TCCfg<BN>;
    constexpr int NBUF = Cfg::NBUF;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    // b_res: the weight planes of this CTA's column block (all k-blocks, hi and lo) stay RESIDENT in shared memory and a stage
    // carries the A tile only -- for the short contractions of the backward (dH = gc W^T: Kc = F) the planes are the same for
    // every row tile, and re-fetching them per tile doubled the L2 -> shared-memory traffic the splitters wait for (ncu: 80 %
    // of the splitter warps' time on `raw_full`)
    const uint32_t stage_bytes = b_res ? 2u * TC_A_BYTES : (uint32_t)Cfg::STAGE_BYTES;
    const int nkb_ = (Kc + TC_BK - 1) / TC_BK;
    uint8_t* bres = smem + (size_t)TC_STAGES * stage_bytes;                           // [nkb][hi | lo] when b_res
    uint64_t* bars = reinterpret_cast<uint64_t*>(bres + (b_res ? (size_t)nkb_ * 2 * Cfg::B_BYTES : 0));
    uint64_t* raw_full = bars;                         // [STAGES]  TMA bytes landed              -> splitters
    uint64_t* full = bars + TC_STAGES;                 // [STAGES]  lo plane written              -> MMA
    uint64_t* empty = bars + 2 * TC_STAGES;            // [STAGES]  MMAs retired (tcgen05.commit) -> TMA
    uint64_t* tfull = bars + 3 * TC_STAGES;            // [NBUF]    accumulator chunk complete    -> epilogue
    uint64_t* tempty = bars + 3 * TC_STAGES + NBUF;    // [NBUF]    accumulator drained           -> MMA
    uint64_t* bfull = bars + 3 * TC_STAGES + 2 * NBUF;   // resident weight planes landed           -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 2 * NBUF + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * BN;
    const int nkb = (Kc + TC_BK - 1) / TC_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(raw_full + s, 1);
            mbar_init(full + s, TC_SPLITTERS / 32);
            mbar_init(empty + s, 1);
        }
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(tfull + b, 1);
            mbar_init(tempty + b, TC_EPILOGUE / 32);
        }
        mbar_init(bfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBlo) : "memory");
    }
    if (warp == 5) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =================================================================== TMA producer (one lane)
        if (lane == 0) {
            uint32_t it = 0;
            if (b_res) {
                mbar_arrive_expect_tx(bfull, (uint32_t)nkb * 2u * Cfg::B_BYTES);
                for (int kb = 0; kb < nkb; ++kb) {
                    tma_load_2d(bres + (size_t)kb * 2 * Cfg::B_BYTES, &mapBhi, bfull, kb * TC_BK, n0);
                    tma_load_2d(bres + (size_t)kb * 2 * Cfg::B_BYTES + Cfg::B_BYTES, &mapBlo, bfull, kb * TC_BK, n0);
                }
            }
            for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x) {
                const int m0 = tile * TC_BM;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    mbar_wait(empty + s, ((it / TC_STAGES) & 1) ^ 1);
                    uint8_t* st = smem + (size_t)s * stage_bytes;
                    mbar_arrive_expect_tx(raw_full + s, b_res ? (uint32_t)TC_A_BYTES : Cfg::TX_BYTES);
                    tma_load_2d(st, &mapA, raw_full + s, kb * TC_BK, m0);                                  // raw A (= hi)
                    if (!b_res) {
                        tma_load_2d(st + 2 * TC_A_BYTES, &mapBhi, raw_full + s, kb * TC_BK, n0);               // Bt_hi
                        tma_load_2d(st + 2 * TC_A_BYTES + Cfg::B_BYTES, &mapBlo, raw_full + s, kb * TC_BK, n0); // Bt_lo
                    }
                }
            }
        }
    } else if (warp <= 4) {
        // =================================================================== splitters: lo = a - trunc_tf32(a)
        const int tid = threadIdx.x - 32;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % TC_STAGES;
                mbar_wait(raw_full + s, (it / TC_STAGES) & 1);
                const uint8_t* a_raw = smem + (size_t)s * stage_bytes;
                uint8_t* a_lo = smem + (size_t)s * stage_bytes + TC_A_BYTES;
#pragma unroll
                for (int j = 0; j < TC_A_BYTES / 16 / TC_SPLITTERS; ++j) {
                    const int off = (tid + TC_SPLITTERS * j) * 16;
                    const float4 v = *reinterpret_cast<const float4*>(a_raw + off);
                    float4 l;
                    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                    *reinterpret_cast<float4*>(a_lo + off) = l;
                }
                fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(full + s);
            }
        }
    } else if (warp == 5) {
        // =================================================================== MMA issuer (one lane)
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32<BN>();
            uint32_t it = 0, cc = 0;
            if (b_res) mbar_wait(bfull, 0);
            for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x) {
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t buf = cc % NBUF;
                    const bool chunk_start = (kb % chunk_kb) == 0;
                    if (chunk_start) {
                        mbar_wait(tempty + buf, ((cc / NBUF) & 1) ^ 1);    // epilogue has drained this accumulator
                        tc_fence_after();
                    }
                    const int s = it % TC_STAGES;
                    mbar_wait(full + s, (it / TC_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t sb = b_res ? smem_u32(bres + (size_t)kb * 2 * Cfg::B_BYTES) : sa + 2 * TC_A_BYTES;
                    const uint64_t da_hi = make_kmajor_sw128_desc(sa);
                    const uint64_t da_lo = make_kmajor_sw128_desc(sa + TC_A_BYTES);
                    const uint64_t db_hi = make_kmajor_sw128_desc(sb);
                    const uint64_t db_lo = make_kmajor_sw128_desc(sb + Cfg::B_BYTES);
                    const uint32_t d = tmem_base + buf * BN;
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);      // 8 TF32 = 32 bytes along K inside the swizzled row
                        umma_tf32(d, da_lo + adv, db_hi + adv, idesc, (chunk_start && k == 0) ? 0u : 1u);
                        umma_tf32(d, da_hi + adv, db_lo + adv, idesc, 1u);
                        umma_tf32(d, da_hi + adv, db_hi + adv, idesc, 1u);
                    }
                    umma_commit(empty + s);                                    // stage reusable once these MMAs retire
                    if ((kb % chunk_kb) == chunk_kb - 1 || kb == nkb - 1) {
                        umma_commit(tfull + buf);                              // chunk complete -> epilogue
                        ++cc;
                    }
                }
            }
        }
    } else {
        // =================================================================== epilogue
        const int quarter = warp & 3;                       // TMEM lanes this warp may touch: [32*quarter, +32)
        const int nchunks = (nkb + chunk_kb - 1) / chunk_kb;
        uint32_t cc = 0;
        for (int tile = blockIdx.x; tile < n_mtiles; tile += gridDim.x) {
            const int64_t row = (int64_t)tile * TC_BM + quarter * 32 + lane;
            float acc[BN];
#pragma unroll
            for (int j = 0; j < BN; ++j) acc[j] = 0.f;
            for (int ch = 0; ch < nchunks; ++ch, ++cc) {
                const uint32_t buf = cc % NBUF;
                mbar_wait(tfull + buf, (cc / NBUF) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * BN;
#pragma unroll
                for (int j0 = 0; j0 < BN; j0 += 32) {
                    float v[32];
                    tmem_ld32(taddr + j0, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[j0 + i] += v[i];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + buf);
            }
            if (row < M) {
                float* dst = C + row * ldc + n0;
                const bool vec = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (n0 % 4 == 0);
#pragma unroll
                for (int j = 0; j < BN; j += 4) {
                    float o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float t = acc[j + i];
                        if (bias && n0 + j + i < Nc) t += __ldg(bias + n0 + j + i);
                        if (epi == GNNML3_EPI_RELU) t = fmaxf(t, 0.f);
                        o[i] = t;
                    }
                    if (vec && n0 + j + 3 < Nc) {
                        *reinterpret_cast<float4*>(dst + j) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (n0 + j + i < Nc) dst[j + i] = o[i];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace gnnml3

using namespace gnnml3;

static inline int tc_bn_for(int Nc) { return Nc <= 32 ? 32 : 64; }

extern "C" int gnnml3_gemm_nn_tc_supported(int64_t lda, int Nc, int Kc) {
    return (lda % 4 == 0 && Nc >= 1 && Kc >= 1 && Nc <= 4096 && Kc <= 8192) ? 1 : 0;
}

extern "C" size_t gnnml3_gemm_nn_tc_workspace_bytes(int Nc, int Kc) {
    const int BN = tc_bn_for(Nc);
    const size_t Npad = (size_t)cdiv(Nc, BN) * BN, Kpad = (size_t)cdiv(Kc, TC_BK) * TC_BK;
    return align_up(2 * Npad * Kpad * sizeof(float), 256);
}

template <int BN>
static int launch_tc(const float* A, int64_t lda, const float* hi, const float* lo, int Kpad, int Npad, const float* bias,
                     float* C, int64_t ldc, int64_t M, int Nc, int Kc, int epi, int chunk_kb, cudaStream_t st) {
    using Cfg = TCCfg<BN>;
    static bool configured[64] = {};
    if (auto once_ = first_use_on_device(configured)) {
        GNNML3_CUDA(cudaFuncSetAttribute(k_gemm_nn_tc<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    }
    CUtensorMap mapA, mapBhi, mapBlo;
    int rc;
    if ((rc = make_map(&mapA, A, M, Kc, lda, TC_BM))) return rc;
    if ((rc = make_map(&mapBhi, hi, Npad, Kpad, Kpad, BN))) return rc;
    if ((rc = make_map(&mapBlo, lo, Npad, Kpad, Kpad, BN))) return rc;
    const int n_mtiles = cdiv(M, TC_BM);
    const int gy = cdiv(Nc, BN);
    int gx = kNumSMs / gy;
    if (gx < 1) gx = 1;
    if (gx > n_mtiles) gx = n_mtiles;
    // resident weight planes when they fit beside the four A stages and every CTA walks more than a few row tiles
    const int nkb = cdiv(Kc, TC_BK);
    const size_t res_smem = (size_t)TC_STAGES * 2 * TC_A_BYTES + (size_t)nkb * 2 * Cfg::B_BYTES + 1024 + 512;
    const int b_res = (res_smem <= Cfg::SMEM && n_mtiles >= 4 * gx) ? 1 : 0;
    k_gemm_nn_tc<BN><<<dim3(gx, gy), TC_THREADS, b_res ? res_smem : Cfg::SMEM, st>>>(mapA, mapBhi, mapBlo, bias, C, ldc, M, Nc, Kc, epi,
                                                                                   chunk_kb, n_mtiles, b_res);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

extern "C" int gnnml3_gemm_nn_tc(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, float* C,
                                 int64_t ldc, int64_t M, int Nc, int Kc, int epilogue, int chunk_kblocks, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
    GNNML3_REQUIRE(M > 0 && Nc > 0 && Kc > 0, "gemm_nn_tc: bad shape");
    GNNML3_REQUIRE(A && B && C && workspace, "gemm_nn_tc: NULL pointer");
    GNNML3_REQUIRE(lda >= Kc && ldb >= Nc && ldc >= Nc, "gemm_nn_tc: leading dimensions too small");
    GNNML3_REQUIRE(gnnml3_gemm_nn_tc_supported(lda, Nc, Kc) && (uintptr_t)A % 16 == 0,
                   "gemm_nn_tc: A rows must be 16-byte aligned (lda %% 4 == 0)");
    GNNML3_REQUIRE(epilogue == GNNML3_EPI_NONE || epilogue == GNNML3_EPI_RELU, "gemm_nn_tc: unknown epilogue");
    GNNML3_REQUIRE(M < (1ll << 31) - TC_BM, "gemm_nn_tc: M too large");
    if (workspace_bytes < gnnml3_gemm_nn_tc_workspace_bytes(Nc, Kc))
        return set_err(GNNML3_ERR_WORKSPACE, "gemm_nn_tc: workspace too small");
    cudaStream_t st = (cudaStream_t)stream_;
    const int BN = tc_bn_for(Nc);
    const int Npad = cdiv(Nc, BN) * BN, Kpad = cdiv(Kc, TC_BK) * TC_BK;
    float* hi = (float*)workspace;
    float* lo = hi + (size_t)Npad * Kpad;
    k_prep_weights_tc<<<cdiv((int64_t)Npad * Kpad, 256) > 592 ? 592 : cdiv((int64_t)Npad * Kpad, 256), 256, 0, st>>>(
        B, ldb, Kc, Nc, Kpad, Npad, hi, lo);
    GNNML3_LAUNCH_CHECK();
    const int chunk = chunk_kblocks > 0 ? chunk_kblocks : 4;   // 4 k-blocks (48 MMAs) per TMEM chunk: rel. error ~1e-6
    if (BN == 32) return launch_tc<32>(A, lda, hi, lo, Kpad, Npad, bias, C, ldc, M, Nc, Kc, epilogue, chunk, st);
    return launch_tc<64>(A, lda, hi, lo, Kpad, Npad, bias, C, ldc, M, Nc, Kc, epilogue, chunk, st);
}
