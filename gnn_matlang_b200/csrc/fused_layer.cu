// Fused aggregate + project kernel of the GNNML3 layer on Blackwell tensor cores.
//
// Reference semantics (libs/spect_conv.py:70-80,93-94 and :208-212):
//     conv[t, :] = sum_k ( sum_{e: dst_e = t} ea[e, k] * x[src_e, :] ) W_k  + bias            (SpectConv)
//     y[t, :]    = [ relu(conv[t, :]) || tanh(x[t] W11^T + b11) * tanh(x[t] W12^T + b12) ]    (ML3Layer)
// The reference materialises K message tensors [E, Fi], K scatter results [N, Fi] and K matmul results; the
// two-kernel design of spmm.cu + gemm*.cu still round-trips H = [P_0(x) .. P_{K-1}(x)]  ([N, K*Fi]) through HBM.
// Here H never leaves the SM: a persistent CTA owns a tile of dst rows,
//   * aggregator warps walk the tile's CSR rows (4 lanes per row, 8 features per lane, KT supports in registers,
//     128-bit gathers of the source rows, summation in edge order, no atomics) and write each 32-column block
//     of H straight into shared memory in the UMMA K-major SWIZZLE_128B layout, as the raw FP32 plane (= the hi
//     part: the tensor core truncates to TF32 itself) and the residual plane lo = a - trunc(a);
//   * one thread issues tcgen05.mma kind::tf32 (lo*hi + hi*lo + hi*hi: FP32-grade 3xTF32) against the pre-split
//     weight planes that a TMA warp streams from L2; accumulators live in tensor memory;
//   * epilogue warps drain TMEM (round-to-nearest chunk sums), add the bias and apply ReLU / tanh*tanh gating,
//     writing y (and the two tanh factors the backward needs) -- or a plain [N, Nc] result.
// The same kernel computes dx in the backward: rows = source nodes over the transposed CSR, gathered matrix =
// d pre (conv columns), weights = W_k^T, and the gate gradients enter as one more k-block ("self" block) that
// accumulates into the same output columns.  SpectConv(selfconn=True) uses that mode in the forward, too.
#include "tc_common.cuh"

namespace gnnml3 {

constexpr int FL_STAGES = 4;
constexpr int FL_PLANE = 128 * 128;   // bytes of one [128 rows x 32 FP32] k-block plane
constexpr int FL_CTRL_WARPS = 6;      // warps 0-3: epilogue | warp 4: TMA (weights) | warp 5: MMA issuer | then aggregators

template <int BN>
struct FLCfg {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = 2 * FL_PLANE + 2 * B_BYTES;     // A raw | A lo | B hi | B lo
    static constexpr int NBUF = 256 / BN;                              // TMEM chunk buffers
    static constexpr int TMEM_COLS = 256;
    static constexpr size_t SMEM = (size_t)FL_STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 512 /*barriers*/;
};

struct FLParams {
    const int* rowptr;      // [N+1] CSR over the rows of this launch
    const int* col;         // [E]   gathered row of X per CSR slot
    const int* eperm;       // [E]   row of `ea` per CSR slot (NULL: identity)
    const float* ea;        // [E, Kstride]
    int Kstride, K;
    const float* X;         // gathered matrix [*, F], row stride ldx (16-byte aligned rows)
    int64_t ldx;
    int F;
    const float* S;         // self block [N, Fs] (row t of the tile itself), NULL if self_mode == 0
    int64_t lds;
    int Fs;
    int self_mode;          // 0 none | 1 own output columns (BNS wide: the ML3 gates) | 2 accumulates into the main columns
    int BNS;                // MMA N of the self block in mode 1 (16 or 32)
    int64_t N;
    int n_tiles;
    int nfh;                // 32-wide feature blocks per support = ceil(F / 32)
    int nkb_main;           // nfh * K
    int chunk_kb;           // k-blocks per TMEM chunk
    const float* bias;      // [Nc] or NULL
    const float* bias_s;    // [2G] or NULL (mode 1)
    float* out;             // plain: [N, Nc]; ml3: y [N, Fo + G]
    int64_t ldo;
    int Nc;                 // main output columns (= Fo)
    float* aux;             // ml3 with G > 0: [N, 2G] tanh factors (t1 | t2)
    int64_t ldaux;
    int G;
    int epi;                // 0 plain (+bias) | 1 ml3: relu on the main columns, gating on the self columns
};

template <int KT>
__device__ __forceinline__ void fl_load_ea(const float* __restrict__ p, float (&w)[KT]) {
    if constexpr (KT % 4 == 0) {
#pragma unroll
        for (int k = 0; k < KT; k += 4) {
            const float4 v = ldg4(p + k);
            w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w;
        }
    } else if constexpr (KT % 2 == 0) {
#pragma unroll
        for (int k = 0; k < KT; k += 2) {
            const float2 v = ldg2(p + k);
            w[k] = v.x; w[k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < KT; ++k) w[k] = __ldg(p + k);
    }
}

__device__ __forceinline__ float fl_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// raw plane + residual plane of one 16-byte chunk
__device__ __forceinline__ void fl_store_chunk(uint8_t* a_raw, uint32_t off, const float* v) {
    *reinterpret_cast<float4*>(a_raw + off) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(a_raw + FL_PLANE + off) = make_float4(fl_lo(v[0]), fl_lo(v[1]), fl_lo(v[2]), fl_lo(v[3]));
}

template <int KT, int BN, int NAGG>
__global__ void __launch_bounds__(32 * (FL_CTRL_WARPS + NAGG), 1)
k_fused_agg_proj(const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
                 const __grid_constant__ CUtensorMap mapShi, const __grid_constant__ CUtensorMap mapSlo,
                 const __grid_constant__ FLParams P) {
    using Cfg = FLCfg<BN>;
    constexpr int NBUF = Cfg::NBUF;
    constexpr int ROWS = 8 * NAGG;                      // dst rows per tile (<= 128 = the MMA's M)
    static_assert(ROWS <= 128, "tile rows exceed the MMA M");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)FL_STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;                             // [STAGES]  A planes written + weight bytes landed -> MMA
    uint64_t* empty = bars + FL_STAGES;                // [STAGES]  MMAs retired                           -> writers
    uint64_t* tfull = bars + 2 * FL_STAGES;            // [NBUF]    accumulator chunk complete             -> epilogue
    uint64_t* tempty = bars + 2 * FL_STAGES + NBUF;    // [NBUF]    accumulator drained                    -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * FL_STAGES + 2 * NBUF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb_chain = P.nkb_main + (P.self_mode == 2 ? 1 : 0);     // k-blocks accumulated into the main columns
    const int nkb_total = P.nkb_main + (P.self_mode != 0 ? 1 : 0);

    // rows ROWS..127 of the A planes are never written: clear them once so that no NaN pattern is ever multiplied
    for (int i = threadIdx.x; i < FL_STAGES * Cfg::STAGE_BYTES / 16; i += blockDim.x)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        for (int s = 0; s < FL_STAGES; ++s) {
            mbar_init(full + s, NAGG + 1);
            mbar_init(empty + s, 1);
        }
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(tfull + b, 1);
            mbar_init(tempty + b, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBlo) : "memory");
    }
    if (warp == 5) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // =================================================================== TMA: weight planes of every k-block
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < nkb_total; ++kb, ++it) {
                    const int s = it % FL_STAGES;
                    mbar_wait(empty + s, ((it / FL_STAGES) & 1) ^ 1);
                    uint8_t* st = smem + (size_t)s * Cfg::STAGE_BYTES + 2 * FL_PLANE;
                    if (kb < nkb_chain) {
                        mbar_arrive_expect_tx(full + s, 2 * Cfg::B_BYTES);
                        tma_load_2d(st, &mapBhi, full + s, 0, kb * BN);
                        tma_load_2d(st + Cfg::B_BYTES, &mapBlo, full + s, 0, kb * BN);
                    } else {
                        mbar_arrive_expect_tx(full + s, 2 * P.BNS * 128);
                        tma_load_2d(st, &mapShi, full + s, 0, 0);
                        tma_load_2d(st + Cfg::B_BYTES, &mapSlo, full + s, 0, 0);
                    }
                }
            }
        }
    } else if (warp == 5) {
        // =================================================================== MMA issuer (one lane)
        if (lane == 0) {
            const uint32_t idesc_main = make_idesc_tf32_mn(128, BN);
            const uint32_t idesc_self = make_idesc_tf32_mn(128, P.BNS > 0 ? P.BNS : 16);
            uint32_t it = 0, cc = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < nkb_total; ++kb, ++it) {
                    const bool is_self = kb >= nkb_chain;
                    const bool chunk_start = is_self || (kb % P.chunk_kb) == 0;
                    const bool chunk_end = is_self || (kb % P.chunk_kb) == P.chunk_kb - 1 || kb == nkb_chain - 1;
                    const uint32_t buf = cc % NBUF;
                    if (chunk_start) {
                        mbar_wait(tempty + buf, ((cc / NBUF) & 1) ^ 1);    // epilogue has drained this accumulator
                        tc_fence_after();
                    }
                    const int s = it % FL_STAGES;
                    mbar_wait(full + s, (it / FL_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)s * Cfg::STAGE_BYTES);
                    const uint64_t da_hi = make_kmajor_sw128_desc(sa);
                    const uint64_t da_lo = make_kmajor_sw128_desc(sa + FL_PLANE);
                    const uint64_t db_hi = make_kmajor_sw128_desc(sa + 2 * FL_PLANE);
                    const uint64_t db_lo = make_kmajor_sw128_desc(sa + 2 * FL_PLANE + Cfg::B_BYTES);
                    const uint32_t d = tmem_base + buf * BN;
                    const uint32_t idesc = is_self ? idesc_self : idesc_main;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);      // 8 TF32 = 32 bytes along K inside the swizzled row
                        umma_tf32(d, da_lo + adv, db_hi + adv, idesc, (chunk_start && k == 0) ? 0u : 1u);
                        umma_tf32(d, da_hi + adv, db_lo + adv, idesc, 1u);
                        umma_tf32(d, da_hi + adv, db_hi + adv, idesc, 1u);
                    }
                    umma_commit(empty + s);                                    // stage reusable once these MMAs retire
                    if (chunk_end) {
                        umma_commit(tfull + buf);                              // chunk complete -> epilogue
                        ++cc;
                    }
                }
            }
        }
    } else if (warp < 4) {
        // =================================================================== epilogue (4 warps, one TMEM lane quarter each)
        const int quarter = warp & 3;
        const int nchunks = (nkb_chain + P.chunk_kb - 1) / P.chunk_kb;
        const bool live_quarter = quarter * 32 < ROWS;
        const int Fo = P.Nc, G = P.G;
        uint32_t cc = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
            const int rloc = quarter * 32 + lane;
            const int64_t row = (int64_t)tile * ROWS + rloc;
            const bool live = live_quarter && rloc < ROWS && row < P.N;
            float acc[BN];
#pragma unroll
            for (int j = 0; j < BN; ++j) acc[j] = 0.f;
            for (int ch = 0; ch < nchunks; ++ch, ++cc) {
                const uint32_t buf = cc % NBUF;
                mbar_wait(tfull + buf, (cc / NBUF) & 1);
                tc_fence_after();
                if (live_quarter) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * BN;
#pragma unroll
                    for (int j0 = 0; j0 < BN; j0 += 32) {
                        float v[32];
                        tmem_ld32(taddr + j0, v);
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[j0 + i] += v[i];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + buf);
            }
            // ---- main columns: bias (+ ReLU) and row store
            if (live) {
                float* dst = P.out + row * P.ldo;
                const bool vec = (P.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0);
#pragma unroll
                for (int j = 0; j < BN; j += 4) {
                    float o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float t = acc[j + i];
                        if (P.bias && j + i < Fo) t += __ldg(P.bias + j + i);
                        if (P.epi == 1) t = fmaxf(t, 0.f);
                        o[i] = t;
                    }
                    if (vec && j + 3 < Fo) {
                        *reinterpret_cast<float4*>(dst + j) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (j + i < Fo) dst[j + i] = o[i];
                    }
                }
            }
            // ---- gate columns (own TMEM chunk): y[:, Fo + j] = tanh(p1_j) * tanh(p2_j); aux = [tanh(p1) | tanh(p2)]
            if (P.self_mode == 1) {
                const uint32_t buf = cc % NBUF;
                mbar_wait(tfull + buf, (cc / NBUF) & 1);
                tc_fence_after();
                float v[32];
                if (live_quarter) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * BN;
                    tmem_ld32(taddr, v);       // BNS <= 32 <= BN columns of this buffer are meaningful
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + buf);
                ++cc;
                if (live) {
                    float* dst = P.out + row * P.ldo + Fo;
                    float* ax = P.aux + row * P.ldaux;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (j < G) {      // the self weight planes interleave the two gates: column 2j = p1_j, 2j+1 = p2_j
                            float p1 = v[2 * j], p2 = v[2 * j + 1];
                            if (P.bias_s) {
                                p1 += __ldg(P.bias_s + j);
                                p2 += __ldg(P.bias_s + G + j);
                            }
                            const float t1 = tanhf(p1), t2 = tanhf(p2);
                            dst[j] = t1 * t2;
                            ax[j] = t1;
                            ax[G + j] = t2;
                        }
                    }
                }
            }
        }
    } else if (warp >= FL_CTRL_WARPS) {
        // =================================================================== aggregators
        const int aw = warp - FL_CTRL_WARPS;
        const int q = lane >> 2, g = lane & 3;
        const int rq = (q >> 1) | ((q & 1) << 2);          // row of the 8-row group owned by this quad: 0,4,1,5,2,6,3,7
        const int rloc = aw * 8 + rq;                      // (bank-conflict-free 128-bit stores into the swizzled plane)
        const uint32_t row_off = (uint32_t)aw * 1024u + (uint32_t)rq * 128u;
        const uint32_t off0 = row_off + (uint32_t)((g ^ rq) << 4);            // chunk g     (features 4g .. 4g+3)
        const uint32_t off1 = row_off + (uint32_t)(((g + 4) ^ rq) << 4);      // chunk g + 4 (features 16+4g .. 16+4g+3)
        const int* __restrict__ rowptr = P.rowptr;
        const int* __restrict__ col = P.col;
        const int* __restrict__ eperm = P.eperm;
        const float* __restrict__ ea = P.ea;
        const float* __restrict__ X = P.X;
        const int64_t ldx = P.ldx;
        const int Kstride = P.Kstride;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
            const int64_t row = (int64_t)tile * ROWS + rloc;
            int rs = 0, re = 0;
            if (row < P.N) {
                rs = __ldg(rowptr + row);
                re = __ldg(rowptr + row + 1);
            }
            for (int fh = 0; fh < P.nfh; ++fh) {
                const int f0 = fh * 32 + g * 4, f1 = f0 + 16;
                const bool v0 = f0 < P.F, v1 = f1 < P.F;
                for (int k0 = 0; k0 < P.K; k0 += KT) {
                    float acc[KT][8];
#pragma unroll
                    for (int k = 0; k < KT; ++k)
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
                    constexpr int U = 2;   // edges in flight per lane
                    for (int p0 = rs; p0 < re; p0 += U) {
                        int sidx[U], eidx[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int p = min(p0 + u, re - 1);
                            sidx[u] = __ldg(col + p);
                            eidx[u] = eperm ? __ldg(eperm + p) : p;
                        }
                        float w[U][KT];
                        float4 xa[U], xb[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            fl_load_ea<KT>(ea + (int64_t)eidx[u] * Kstride + k0, w[u]);
                            const float* xr = X + (int64_t)sidx[u] * ldx;
                            xa[u] = v0 ? ldg4(xr + f0) : make_float4(0.f, 0.f, 0.f, 0.f);
                            xb[u] = v1 ? ldg4(xr + f1) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if (p0 + u < re) {       // edges of a row in their original order (the reference's CPU order)
#pragma unroll
                                for (int k = 0; k < KT; ++k) {
                                    acc[k][0] = fmaf(w[u][k], xa[u].x, acc[k][0]);
                                    acc[k][1] = fmaf(w[u][k], xa[u].y, acc[k][1]);
                                    acc[k][2] = fmaf(w[u][k], xa[u].z, acc[k][2]);
                                    acc[k][3] = fmaf(w[u][k], xa[u].w, acc[k][3]);
                                    acc[k][4] = fmaf(w[u][k], xb[u].x, acc[k][4]);
                                    acc[k][5] = fmaf(w[u][k], xb[u].y, acc[k][5]);
                                    acc[k][6] = fmaf(w[u][k], xb[u].z, acc[k][6]);
                                    acc[k][7] = fmaf(w[u][k], xb[u].w, acc[k][7]);
                                }
                            }
                        }
                    }
                    // hand the KT finished k-blocks to the tensor core
#pragma unroll
                    for (int k = 0; k < KT; ++k, ++it) {
                        const int s = it % FL_STAGES;
                        mbar_wait(empty + s, ((it / FL_STAGES) & 1) ^ 1);
                        uint8_t* a_raw = smem + (size_t)s * Cfg::STAGE_BYTES;
                        fl_store_chunk(a_raw, off0, &acc[k][0]);
                        fl_store_chunk(a_raw, off1, &acc[k][4]);
                        fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core (async proxy)
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full + s);
                    }
                }
            }
            if (P.self_mode != 0) {
                float sv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) sv[i] = 0.f;
                if (row < P.N) {
                    const float* sr = P.S + row * P.lds;
                    if (g * 4 < P.Fs) {
                        const float4 t = ldg4(sr + g * 4);
                        sv[0] = t.x; sv[1] = t.y; sv[2] = t.z; sv[3] = t.w;
                    }
                    if (16 + g * 4 < P.Fs) {
                        const float4 t = ldg4(sr + 16 + g * 4);
                        sv[4] = t.x; sv[5] = t.y; sv[6] = t.z; sv[7] = t.w;
                    }
                }
                const int s = it % FL_STAGES;
                mbar_wait(empty + s, ((it / FL_STAGES) & 1) ^ 1);
                uint8_t* a_raw = smem + (size_t)s * Cfg::STAGE_BYTES;
                fl_store_chunk(a_raw, off0, &sv[0]);
                fl_store_chunk(a_raw, off1, &sv[4]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full + s);
                ++it;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// Pre-split (hi = RN TF32, lo = residual), transposed (K-major) weight planes, one [BN x 32] plane per k-block in the
// order the aggregators produce them: kb = fh * K + k covers rows (k * F + fh * 32 + c), c < 32, of Bmain [K*F, Nc];
// kb = nfh * K is the self block Bself [Fs, Nc] when it accumulates into the main columns (mode 2).  In mode 1 the
// self block has its own [BNS x 32] plane pair (Bself [Fs, Ns = 2G], output columns interleaved p1_0 p2_0 p1_1 ..).
__global__ void k_fl_prep_weights(const float* __restrict__ Bmain, int64_t ldb, int K, int F, int Nc, int nfh, int BN,
                                  const float* __restrict__ Bself, int64_t ldbs, int Fs, int Ns, int self_mode, int BNS,
                                  float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ shi,
                                  float* __restrict__ slo) {
    const int nkb_main = nfh * K;
    const int nkb_chain = nkb_main + (self_mode == 2 ? 1 : 0);
    const int total_main = nkb_chain * BN * 32;
    const int total = total_main + (self_mode == 1 ? BNS * 32 : 0);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        float *ph, *pl;
        if (i < total_main) {
            const int kb = i / (BN * 32), n = (i / 32) % BN, c = i % 32;
            if (kb < nkb_main) {
                const int fh = kb / K, k = kb % K, f = fh * 32 + c;
                if (f < F && n < Nc) v = __ldg(Bmain + ((int64_t)k * F + f) * ldb + n);
            } else if (c < Fs && n < Nc) {
                v = __ldg(Bself + (int64_t)c * ldbs + n);
            }
            ph = hi + i;
            pl = lo + i;
        } else {
            const int j = i - total_main;
            const int n = j / 32, c = j % 32;
            const int src_n = (n & 1) ? (Ns / 2 + (n >> 1)) : (n >> 1);      // interleave [p1 | p2] -> p1_0 p2_0 p1_1 p2_1 ..
            if (c < Fs && n < Ns) v = __ldg(Bself + (int64_t)c * ldbs + src_n);
            ph = shi + j;
            pl = slo + j;
        }
        const float h = tf32_rn(v);
        *ph = h;
        *pl = v - h;
    }
}

}  // namespace gnnml3

using namespace gnnml3;

static inline int fl_bn_for(int Nc) { return Nc <= 32 ? 32 : 64; }
static inline int fl_kt_for(int K) {
    if (K >= 4 && K <= 8) return K;
    if (K > 8 && K <= 16 && K % 2 == 0) return K / 2;
    return 0;
}

static int g_fl_nagg16 = [] {
    const char* e = getenv("GNNML3_FUSED_NAGG16");
    return (e && e[0] == '1') ? 1 : 0;
}();

extern "C" int gnnml3_fused_supported(int K, int Kstride, int F, int Nc, int Fs, int self_mode, int Ns) {
    const int kt = fl_kt_for(K);
    if (kt == 0 || F < 1 || F > 256 || Nc < 1 || Nc > 64) return 0;
    if (kt % 4 == 0 && Kstride % 4 != 0) return 0;        // 128-bit edge-weight loads
    if (kt % 4 != 0 && kt % 2 == 0 && Kstride % 2 != 0) return 0;
    if (self_mode != 0 && (Fs < 1 || Fs > 32)) return 0;
    if (self_mode == 1 && (Ns < 1 || Ns > 32)) return 0;
    return 1;
}

extern "C" size_t gnnml3_fused_workspace_bytes(int K, int F, int Nc, int self_mode) {
    const int BN = fl_bn_for(Nc);
    const int nfh = cdiv(F, 32);
    const size_t planes = (size_t)(nfh * K + (self_mode == 2 ? 1 : 0)) * BN * 32;
    return align_up((2 * planes + 2 * 32 * 32) * sizeof(float), 256);
}

template <int KT, int BN, int NAGG>
static int fl_launch(const CUtensorMap& mBhi, const CUtensorMap& mBlo, const CUtensorMap& mShi, const CUtensorMap& mSlo,
                     FLParams& P, cudaStream_t st) {
    using Cfg = FLCfg<BN>;
    static bool configured = false;
    if (!configured) {
        GNNML3_CUDA(cudaFuncSetAttribute(k_fused_agg_proj<KT, BN, NAGG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM));
        configured = true;
    }
    constexpr int ROWS = 8 * NAGG;
    P.n_tiles = cdiv(P.N, ROWS);
    const int grid = P.n_tiles < kNumSMs ? P.n_tiles : kNumSMs;
    k_fused_agg_proj<KT, BN, NAGG><<<grid, 32 * (FL_CTRL_WARPS + NAGG), Cfg::SMEM, st>>>(mBhi, mBlo, mShi, mSlo, P);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

// aggregator warps per CTA by register need: KT x 8 accumulators per lane.  512 threads leave 128 registers per thread
// (KT >= 7), 640 threads 96 (KT <= 6); the experimental 16-warp / 4-supports-per-pass variant runs at 80.
#define FL_DISPATCH_KT(KTV, BNV)                                                                      \
    switch (KTV) {                                                                                    \
        case 4: return fl_launch<4, BNV, 14>(mBhi, mBlo, mShi, mSlo, P, st);                          \
        case 5: return fl_launch<5, BNV, 14>(mBhi, mBlo, mShi, mSlo, P, st);                          \
        case 6: return fl_launch<6, BNV, 14>(mBhi, mBlo, mShi, mSlo, P, st);                          \
        case 7: return fl_launch<7, BNV, 10>(mBhi, mBlo, mShi, mSlo, P, st);                          \
        case 8: return fl_launch<8, BNV, 10>(mBhi, mBlo, mShi, mSlo, P, st);                          \
        default: return set_err(GNNML3_ERR_INVALID, "fused_agg_proj: no kernel for K tile %d", KTV);  \
    }

extern "C" int gnnml3_fused_agg_proj(const int32_t* rowptr, const int32_t* col, const int32_t* eperm, const float* ea,
                                     int Kstride, int K, const float* X, int64_t ldx, int F, const float* S, int64_t lds,
                                     int Fs, int self_mode, const float* Bmain, int64_t ldb, const float* Bself,
                                     int64_t ldbs, int Ns, const float* bias, const float* bias_s, int64_t N, int Nc,
                                     float* out, int64_t ldo, float* aux, int64_t ldaux, int G, int epilogue,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
    GNNML3_REQUIRE(N > 0 && N < (1ll << 31) - 256, "fused_agg_proj: bad N");
    GNNML3_REQUIRE(rowptr && col && ea && X && Bmain && out && workspace, "fused_agg_proj: NULL pointer");
    GNNML3_REQUIRE(gnnml3_fused_supported(K, Kstride, F, Nc, Fs, self_mode, Ns), "fused_agg_proj: unsupported shape "
                   "K=%d Kstride=%d F=%d Nc=%d Fs=%d self_mode=%d Ns=%d", K, Kstride, F, Nc, Fs, self_mode, Ns);
    GNNML3_REQUIRE(ldx % 4 == 0 && (uintptr_t)X % 16 == 0, "fused_agg_proj: X rows must be 16-byte aligned (ldx %% 4 == 0)");
    GNNML3_REQUIRE((uintptr_t)ea % 16 == 0, "fused_agg_proj: ea must be 16-byte aligned");
    GNNML3_REQUIRE(self_mode == 0 || (S && Bself && lds % 4 == 0 && (uintptr_t)S % 16 == 0),
                   "fused_agg_proj: self block needs S, Bself and 16-byte aligned rows");
    GNNML3_REQUIRE(epilogue == 0 || epilogue == 1, "fused_agg_proj: unknown epilogue");
    GNNML3_REQUIRE(self_mode != 1 || (epilogue == 1 && aux && G >= 1 && G <= 16 && Ns == 2 * G),
                   "fused_agg_proj: gate columns need the ML3 epilogue, aux and Ns == 2G <= 32");
    GNNML3_REQUIRE(ldo >= Nc + (self_mode == 1 ? G : 0), "fused_agg_proj: ldo too small");
    if (workspace_bytes < gnnml3_fused_workspace_bytes(K, F, Nc, self_mode))
        return set_err(GNNML3_ERR_WORKSPACE, "fused_agg_proj: workspace too small");
    cudaStream_t st = (cudaStream_t)stream_;
    const int BN = fl_bn_for(Nc);
    const int KT = fl_kt_for(K);
    const int nfh = cdiv(F, 32);
    const int nkb_main = nfh * K;
    const int nkb_chain = nkb_main + (self_mode == 2 ? 1 : 0);
    const int BNS = self_mode == 1 ? (Ns <= 16 ? 16 : 32) : 0;
    float* hi = (float*)workspace;
    float* lo = hi + (size_t)nkb_chain * BN * 32;
    float* shi = lo + (size_t)nkb_chain * BN * 32;
    float* slo = shi + 32 * 32;
    {
        const int total = nkb_chain * BN * 32 + BNS * 32;
        const int blocks = cdiv(total, 256) > 592 ? 592 : cdiv(total, 256);
        k_fl_prep_weights<<<blocks, 256, 0, st>>>(Bmain, ldb, K, F, Nc, nfh, BN, Bself, ldbs, Fs, Ns, self_mode, BNS, hi, lo,
                                                  shi, slo);
        GNNML3_LAUNCH_CHECK();
    }
    CUtensorMap mBhi, mBlo, mShi, mSlo;
    int rc;
    if ((rc = make_map(&mBhi, hi, (int64_t)nkb_chain * BN, 32, 32, BN))) return rc;
    if ((rc = make_map(&mBlo, lo, (int64_t)nkb_chain * BN, 32, 32, BN))) return rc;
    if (self_mode == 1) {
        if ((rc = make_map(&mShi, shi, BNS, 32, 32, BNS))) return rc;
        if ((rc = make_map(&mSlo, slo, BNS, 32, 32, BNS))) return rc;
    } else {
        mShi = mBhi;
        mSlo = mBlo;
    }
    FLParams P;
    P.rowptr = rowptr; P.col = col; P.eperm = eperm; P.ea = ea; P.Kstride = Kstride; P.K = K;
    P.X = X; P.ldx = ldx; P.F = F; P.S = S; P.lds = lds; P.Fs = Fs; P.self_mode = self_mode; P.BNS = BNS;
    P.N = N; P.n_tiles = 0; P.nfh = nfh; P.nkb_main = nkb_main; P.chunk_kb = 4;
    P.bias = bias; P.bias_s = bias_s; P.out = out; P.ldo = ldo; P.Nc = Nc; P.aux = aux; P.ldaux = ldaux; P.G = G;
    P.epi = epilogue;
    if (g_fl_nagg16 && K % 4 == 0 && Kstride % 4 == 0) {     // experiment: 16 aggregator warps, 4 supports per pass
        if (BN == 32) return fl_launch<4, 32, 16>(mBhi, mBlo, mShi, mSlo, P, st);
        return fl_launch<4, 64, 16>(mBhi, mBlo, mShi, mSlo, P, st);
    }
    if (BN == 32) { FL_DISPATCH_KT(KT, 32) }
    FL_DISPATCH_KT(KT, 64)
}
