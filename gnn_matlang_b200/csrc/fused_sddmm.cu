// Fused edge-feature gradient of SpectConv on Blackwell tensor cores:
//     d ea[e, k] = < x[src_e, :],  (grad_out[dst_e, :] W_k^T) >            (reference: autograd through
//     libs/spect_conv.py:76-80 -- the gradient of  sum_k scatter_add(ea[:, k] * x[src]) W_k  w.r.t. edge_attr)
// The two-kernel path first materialises dH = grad_out [W_0^T .. W_{K-1}^T]  ([N, K*Fi], a GEMM) in HBM and then runs an
// SDDMM over the CSR that reads it back.  Here a persistent CTA owns a tile of 64 dst rows:
//   * the tile's rows of grad_out (conv columns of d pre) are written as raw + residual TF32 planes; one thread issues
//     tcgen05.mma with the weights on the M side -- accumulator p covers the 128 (k, i) pairs 128p .. 128p+127 of
//     Wr[(k,i), o] = W[k, i, o]; the hi and lo parts of the weights are CONCATENATED ALONG K (two resident 128-row planes
//     per accumulator, the grad_out planes are simply read twice), so hi and lo weight products add inside the tensor
//     core and every accumulator lane is one finished (k, i) row -- and the tile rows on the N side ([hi | lo], N = 128):
//     16 instructions per tile produce the whole 3xTF32 dH tile in tensor memory (two buffers: the MMAs of the next tile
//     overlap the drain of this one);
//   * the four epilogue warps add the hi/lo column halves and transpose dH[(k,i), t] -> dH[t][(k,i)] into shared memory,
//     each on its own 32 lanes, no exchange between them;
//   * SDDMM warps (4 lanes per row) keep the row's K x 8 slice of dH in registers, read each source row once per edge
//     with 128-bit loads and reduce-scatter the K dot products over the row's lanes (no atomics, fixed order).
//   * (round 2) the source rows come from SHARED MEMORY: in a batched disjoint graph the sources of a tile lie in a window of
//     consecutive rows of X around it (the per-tile windows of the graph plan, gnnml3_tile_windows, shared with the fused layer
//     kernel); a stager warp brings the window in with one TMA box load and copies the tile's CSR slots, one tile ahead, double
//     buffered.  The dependent chain  col -> x row  per edge is two shared-memory loads instead of two L2 round trips (the
//     kernel was bound by exactly that latency: 8 SDDMM warps per SM cannot hide ~1200 cycles per edge pair).  Tiles whose
//     window or slot count does not fit gather from global memory as before -- same arithmetic, same order.
// dH never touches HBM.  Supported: K * 32 <= 256 (K <= 8), Fi <= 32, Fo <= 32 -- every GNNML3 layer of the ZINC
// configuration; other shapes keep the two-kernel path (gnnml3_gemm_nn + gnnml3_sddmm_k).
#include "tc_common.cuh"

namespace gnnml3 {

constexpr int SD_ROWS = 64;                    // dst rows per tile
constexpr int SD_NW = 8;                       // SDDMM warps (8 rows each)
constexpr int SD_CTRL = 6;                     // warps 0-3 epilogue | 4 TMA (weights) + stager | 5 MMA
constexpr int SD_THREADS = 32 * (SD_CTRL + SD_NW);
constexpr int SD_WPLANE = 128 * 128;           // one weight plane: 64 hi + 64 lo rows x 32 FP32
constexpr int SD_GPLANES = 2 * SD_ROWS * 128;  // raw + lo plane of the grad_out tile
constexpr int SD_LD = 260;                     // row stride (floats) of the dH tile in shared memory
constexpr int SD_WIN_ROWS = 256;               // rows of the staged source window (one TMA box; windows belong to 128-row tiles)
constexpr int SD_CAP = 1024;                   // CSR slots of a tile the staging buffer holds
constexpr int SD_STAGE = SD_WIN_ROWS * 128 + SD_CAP * 4 + 1024;   // x window | slot -> source row | {first slot, first window row, staged}; a multiple of 1024 (TMA 128B-swizzle destination)
constexpr size_t SD_SMEM = 4 * SD_WPLANE + SD_GPLANES + (size_t)SD_ROWS * SD_LD * 4 + 2 * (size_t)SD_STAGE + 1024 + 256;

struct SDParams {
    const int* rowptr;
    const int* col;
    const float* X;        // [N, Fi] gather source
    int64_t ldx;
    int Fi;
    const float* GC;       // [N, Fo] grad_out (conv columns of d pre), 16-byte aligned rows
    int64_t ldg;
    int Fo;
    int K;
    int64_t N;
    int n_tiles;
    float* dea;            // [E, K] in CSR slot order
    const int2* tilewin;   // [ceil(N / 128)] {first, last + 1} source row of the 128-row tile's slots; NULL: gather from global memory
    int win_rows;          // rows of the TMA box (<= SD_WIN_ROWS)
};

__device__ __forceinline__ float sd_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

template <int K>
__global__ void __launch_bounds__(SD_THREADS, 1)
k_fused_sddmm(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX, const __grid_constant__ SDParams P) {
    constexpr int NACC = (K * 32 + 127) / 128;       // TMEM accumulators of 128 (k, i) rows each (<= 2)
    constexpr int NP = 2 * NACC;                     // weight planes: hi and lo of every accumulator's rows
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* wres = smem;                                             // [4][SD_WPLANE]
    uint8_t* gpl = smem + 4 * SD_WPLANE;                              // grad_out planes: raw [64 x 128 B] | lo
    uint8_t* stg = gpl + SD_GPLANES;                                  // [2][SD_STAGE] (1024-aligned: TMA destination)
    float* dhs = reinterpret_cast<float*>(stg + 2 * SD_STAGE);        // [SD_ROWS][SD_LD]: the SDDMM warps copy their slices out at once
    uint64_t* bars = reinterpret_cast<uint64_t*>(dhs + SD_ROWS * SD_LD);
    uint64_t* wfull = bars;          // weights landed
    uint64_t* gfull = bars + 1;      // grad_out planes written        (SD_NW arrivals)   -> MMA
    uint64_t* gempty = bars + 2;     // ... consumed                   (MMA commit)       -> SDDMM warps
    uint64_t* tfull = bars + 3;      // [2] dH tile complete in TMEM   (MMA commit)       -> epilogue
    uint64_t* tempty = bars + 5;     // [2] TMEM buffer drained        (4 arrivals)       -> MMA
    uint64_t* hfull = bars + 7;      // dH tile in shared memory       (4 arrivals)       -> SDDMM warps
    uint64_t* hempty = bars + 8;     // ... copied to registers        (SD_NW arrivals)   -> epilogue
    uint64_t* sfull = bars + 9;      // [2] staging buffer complete    (expect_tx + stager arrive) -> SDDMM warps
    uint64_t* sempty = bars + 11;    // [2] ... no longer read         (SD_NW arrivals)   -> stager
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(wfull, 1);
        mbar_init(gfull, SD_NW);
        mbar_init(gempty, 1);
        mbar_init(hfull, 4);
        mbar_init(hempty, SD_NW);
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull + b, 1);
            mbar_init(tempty + b, 4);
            mbar_init(sfull + b, 2);
            mbar_init(sempty + b, SD_NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    }
    if (warp == 5) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(wfull, NP * SD_WPLANE);
            for (int p = 0; p < NP; ++p) tma_load_2d(wres + (size_t)p * SD_WPLANE, &mapW, wfull, 0, p * 128);
        }
        // =================================================================== stager, one tile ahead of the SDDMM warps
        uint32_t tt = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++tt) {
            const uint32_t b = tt & 1;
            uint8_t* sb = stg + (size_t)b * SD_STAGE;
            int* scol = reinterpret_cast<int*>(sb + SD_WIN_ROWS * 128);
            int* smeta = scol + SD_CAP;
            mbar_wait(sempty + b, ((tt >> 1) & 1) ^ 1);
            const int64_t r0 = (int64_t)tile * SD_ROWS;
            const int e0 = __ldg(P.rowptr + r0);
            const int e1 = __ldg(P.rowptr + (r0 + SD_ROWS < P.N ? r0 + SD_ROWS : P.N));
            int2 w = make_int2(0, 0);
            if (P.tilewin) w = __ldg(P.tilewin + (tile >> 1));        // window of the 128-row tile this 64-row tile is half of
            const bool staged = w.y > w.x && w.y - w.x <= P.win_rows && e1 - e0 <= SD_CAP;
            if (lane == 0) {
                smeta[0] = e0; smeta[1] = w.x; smeta[2] = staged ? 1 : 0;
                mbar_arrive_expect_tx(sfull + b, staged ? (uint32_t)P.win_rows * 128u : 0u);
                if (staged) tma_load_2d(sb, &mapX, sfull + b, 0, w.x);
            }
            if (staged)
                for (int e = e0 + lane; e < e1; e += 32) scol[e - e0] = __ldg(P.col + e);
            __syncwarp();
            if (lane == 0) mbar_arrive(sfull + b);
        }
    } else if (warp == 5) {
        // =================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32_mn(128, 2 * SD_ROWS);
            mbar_wait(wfull, 0);
            uint32_t tt = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++tt) {
                const uint32_t tb = tt & 1;
                mbar_wait(gfull, tt & 1);
                mbar_wait(tempty + tb, ((tt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint64_t dg = make_kmajor_sw128_desc(smem_u32(gpl));               // N side: [gc raw ; gc lo] rows
#pragma unroll
                for (int p = 0; p < NACC; ++p) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {                                          // weight hi plane, then lo plane
                        const uint64_t dw = make_kmajor_sw128_desc(smem_u32(wres + (size_t)(2 * p + h) * SD_WPLANE));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);
                            umma_tf32(tmem_base + tb * 256 + p * 128, dw + adv, dg + adv, idesc, (h == 0 && k == 0) ? 0u : 1u);
                        }
                    }
                }
                umma_commit(gempty);
                umma_commit(tfull + tb);
            }
        }
    } else if (warp < 4) {
        // =================================================================== epilogue: TMEM -> dH[t][(k,i)] in shared memory
        // accumulator p: lane m = (k,i) row 128p + m, column n = tile row (n < 64: x gc_hi, n >= 64: x gc_lo)
        const int q = warp;
        uint32_t tt = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++tt) {
            const uint32_t hb = tt & 1;
            mbar_wait(tfull + hb, (tt >> 1) & 1);
            mbar_wait(hempty, (tt & 1) ^ 1);
            tc_fence_after();
            float* dh = dhs;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + hb * 256;
#pragma unroll 1
            for (int p = 0; p < NACC; ++p) {
                float* dcol = dh + 128 * p + 32 * q + lane;
#pragma unroll 1
                for (int t0 = 0; t0 < SD_ROWS; t0 += 16) {
                    float vh[16], vl[16];
                    tmem_ld16(taddr + p * 128 + t0, vh);
                    tmem_ld16(taddr + p * 128 + SD_ROWS + t0, vl);
#pragma unroll
                    for (int i = 0; i < 16; ++i) dcol[(t0 + i) * SD_LD] = vh[i] + vl[i];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(tempty + hb);
                mbar_arrive(hfull);
            }
        }
    } else {
        // =================================================================== SDDMM warps: 8 rows per warp, 4 lanes per row
        const int aw = warp - SD_CTRL;
        const int qd = lane >> 2, g = lane & 3;
        const int rq = (qd >> 1) | ((qd & 1) << 2);            // conflict-free plane stores (see fused_layer.cu)
        const int rloc = aw * 8 + rq;
        const uint32_t row_off = (uint32_t)aw * 1024u + (uint32_t)rq * 128u;
        const uint32_t off0 = row_off + (uint32_t)((g ^ rq) << 4);
        const uint32_t off1 = row_off + (uint32_t)(((g + 4) ^ rq) << 4);
        // where this lane's fully reduced values land after the reduce-scatter over the quad: 2 channels per lane
        constexpr int KP = 8;
        auto write_gc = [&](int tile, uint32_t seq) {
            const int64_t row = (int64_t)tile * SD_ROWS + rloc;
            float a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = 0.f;
            if (row < P.N) {
                const float* gr = P.GC + row * P.ldg;
                if (g * 4 < P.Fo) {
                    const float4 t = ldg4(gr + g * 4);
                    a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
                }
                if (16 + g * 4 < P.Fo) {
                    const float4 t = ldg4(gr + 16 + g * 4);
                    a[4] = t.x; a[5] = t.y; a[6] = t.z; a[7] = t.w;
                }
            }
            mbar_wait(gempty, (seq & 1) ^ 1);
            *reinterpret_cast<float4*>(gpl + off0) = make_float4(a[0], a[1], a[2], a[3]);
            *reinterpret_cast<float4*>(gpl + off1) = make_float4(a[4], a[5], a[6], a[7]);
            *reinterpret_cast<float4*>(gpl + SD_ROWS * 128 + off0) = make_float4(sd_lo(a[0]), sd_lo(a[1]), sd_lo(a[2]), sd_lo(a[3]));
            *reinterpret_cast<float4*>(gpl + SD_ROWS * 128 + off1) = make_float4(sd_lo(a[4]), sd_lo(a[5]), sd_lo(a[6]), sd_lo(a[7]));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(gfull);
        };
        uint32_t tt = 0;
        if ((int)blockIdx.x < P.n_tiles) write_gc(blockIdx.x, 0);
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++tt) {
            if (tile + (int)gridDim.x < P.n_tiles) write_gc(tile + gridDim.x, tt + 1);      // next tile's planes -> MMA runs ahead
            const uint32_t sbi = tt & 1;
            const int64_t row = (int64_t)tile * SD_ROWS + rloc;
            int rs = 0, cnt = 0;
            if (row < P.N) {
                rs = __ldg(P.rowptr + row);
                cnt = __ldg(P.rowptr + row + 1) - rs;
            }
            int maxcnt = cnt;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, o));
            mbar_wait(hfull, tt & 1);
            // the row's slice of dH: channels k = 0..K-1, features 8g .. 8g+7
            const float* dh = dhs + (size_t)rloc * SD_LD + 8 * g;
            float gv[K][8];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(dh + 32 * k);
                const float4 b = *reinterpret_cast<const float4*>(dh + 32 * k + 4);
                gv[k][0] = a.x; gv[k][1] = a.y; gv[k][2] = a.z; gv[k][3] = a.w;
                gv[k][4] = b.x; gv[k][5] = b.y; gv[k][6] = b.z; gv[k][7] = b.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(hempty);               // dH slice is in registers: the buffer may be refilled
            const bool v0 = 8 * g < P.Fi, v1 = 8 * g + 4 < P.Fi;
            // staged operands of this tile
            mbar_wait(sfull + sbi, (tt >> 1) & 1);
            const uint8_t* sb = stg + (size_t)sbi * SD_STAGE;
            const int* scol = reinterpret_cast<const int*>(sb + SD_WIN_ROWS * 128);
            const int e0 = scol[SD_CAP], win0 = scol[SD_CAP + 1];
            const bool staged = scol[SD_CAP + 2] != 0;
            const uint32_t x32 = smem_u32(sb);
            // two edges of the row in flight per iteration (index -> gather chain of the second overlaps the first)
            for (int i = 0; i < maxcnt; i += 2) {
                float part[2][KP];
                bool act[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    act[u] = i + u < cnt;
#pragma unroll
                    for (int k = 0; k < KP; ++k) part[u][k] = 0.f;
                }
                float4 xa[2], xb[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    xa[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    xb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (act[u]) {
                        if (staged) {           // slot -> source row -> window row (128B-swizzled TMA box; columns >= Fi are zero-filled)
                            const int sw = scol[rs - e0 + i + u] - win0;
                            const uint32_t ra = x32 + (uint32_t)sw * 128u;
                            const uint32_t sx = (uint32_t)(sw & 7);
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(xa[u].x), "=f"(xa[u].y), "=f"(xa[u].z), "=f"(xa[u].w) : "r"(ra + (((2u * g) ^ sx) << 4)));
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(xb[u].x), "=f"(xb[u].y), "=f"(xb[u].z), "=f"(xb[u].w) : "r"(ra + (((2u * g + 1u) ^ sx) << 4)));
                        } else {
                            const int s = __ldg(P.col + rs + i + u);
                            const float* xr = P.X + (int64_t)s * P.ldx + 8 * g;
                            if (v0) xa[u] = ldg4(xr);
                            if (v1) xb[u] = ldg4(xr + 4);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        float t = gv[k][0] * xa[u].x;
                        t = fmaf(gv[k][1], xa[u].y, t);
                        t = fmaf(gv[k][2], xa[u].z, t);
                        t = fmaf(gv[k][3], xa[u].w, t);
                        t = fmaf(gv[k][4], xb[u].x, t);
                        t = fmaf(gv[k][5], xb[u].y, t);
                        t = fmaf(gv[k][6], xb[u].z, t);
                        t = fmaf(gv[k][7], xb[u].w, t);
                        part[u][k] = t;
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    // reduce-scatter over the 4 lanes of the row: 8 -> 4 -> 2 values per lane
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const bool upper = (g & 2) != 0;
                        const float keep = upper ? part[u][q4 + 4] : part[u][q4];
                        const float send = upper ? part[u][q4] : part[u][q4 + 4];
                        part[u][q4] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                    }
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) {
                        const bool upper = (g & 1) != 0;
                        const float keep = upper ? part[u][q2 + 2] : part[u][q2];
                        const float send = upper ? part[u][q2] : part[u][q2 + 2];
                        part[u][q2] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                    }
                    // lane g now owns channels kbase, kbase + 1 with kbase = 4 * (g >> 1) + 2 * (g & 1)
                    if (act[u]) {
                        const int kbase = 4 * (g >> 1) + 2 * (g & 1);
                        float* o = P.dea + (int64_t)(rs + i + u) * K + kbase;
                        if (kbase + 1 < K) {
                            *reinterpret_cast<float2*>(o) = make_float2(part[u][0], part[u][1]);
                        } else if (kbase < K) {
                            o[0] = part[u][0];
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sempty + sbi);         // staging buffer no longer read by this warp
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// Weight planes: plane 2a + h (a = accumulator, h = 0 hi / 1 lo), row r: Wr[(k, i) = 128a + r][o] = W[k, i, o]
// (i < Fi, o < Fo, zero padded; hi = RN-TF32, lo = residual)
__global__ void k_sd_prep_weights(const float* __restrict__ W, int K, int Fi, int Fo, float* __restrict__ planes) {
    const int nacc = (K * 32 + 127) / 128;
    const int total = 2 * nacc * 128 * 32;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int pl = idx / (128 * 32), r = (idx / 32) % 128, o = idx % 32;
        const int ki = 128 * (pl >> 1) + r, k = ki / 32, i = ki % 32;
        float v = 0.f;
        if (k < K && i < Fi && o < Fo) v = __ldg(W + ((int64_t)k * Fi + i) * Fo + o);
        const float h = tf32_rn(v);
        planes[idx] = (pl & 1) ? v - h : h;
    }
}

}  // namespace gnnml3

using namespace gnnml3;

extern "C" int gnnml3_fused_sddmm_supported(int K, int Fi, int Fo) {
    return (K >= 2 && K <= 8 && K % 2 == 0 && Fi >= 1 && Fi <= 32 && Fo >= 1 && Fo <= 32) ? 1 : 0;
}

extern "C" size_t gnnml3_fused_sddmm_workspace_bytes(int K) { return align_up((size_t)4 * SD_WPLANE, 256); }

template <int K>
static int sd_launch(const CUtensorMap& mW, const CUtensorMap& mX, SDParams& P, cudaStream_t st) {
    static bool configured[64] = {};
    if (auto once_ = first_use_on_device(configured)) {
        GNNML3_CUDA(cudaFuncSetAttribute(k_fused_sddmm<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SD_SMEM));
    }
    const int grid = P.n_tiles < kNumSMs ? P.n_tiles : kNumSMs;
    k_fused_sddmm<K><<<grid, SD_THREADS, SD_SMEM, st>>>(mW, mX, P);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

namespace gnnml3 {
int ts_make_xmap(CUtensorMap* map, const float* base, int64_t rows, int F, int64_t ld, int box_rows);   // fused_layer_ts.cu
}

extern "C" int gnnml3_fused_sddmm(const int32_t* rowptr, const int32_t* col, const int32_t* win, const float* X, int64_t ldx, int Fi,
                                  const float* GC, int64_t ldg, int Fo, const float* W, int K, int64_t N, float* dea, void* workspace,
                                  size_t workspace_bytes, void* stream_) {
    GNNML3_REQUIRE(N > 0 && N < (1ll << 31) - 256, "fused_sddmm: bad N");
    GNNML3_REQUIRE(rowptr && col && X && GC && W && dea && workspace, "fused_sddmm: NULL pointer");
    GNNML3_REQUIRE(gnnml3_fused_sddmm_supported(K, Fi, Fo), "fused_sddmm: unsupported shape K=%d Fi=%d Fo=%d", K, Fi, Fo);
    GNNML3_REQUIRE(ldx % 4 == 0 && (uintptr_t)X % 16 == 0 && ldg % 4 == 0 && (uintptr_t)GC % 16 == 0,
                   "fused_sddmm: X and GC rows must be 16-byte aligned");
    GNNML3_REQUIRE((uintptr_t)dea % 8 == 0, "fused_sddmm: dea must be 8-byte aligned");
    if (workspace_bytes < gnnml3_fused_sddmm_workspace_bytes(K)) return set_err(GNNML3_ERR_WORKSPACE, "fused_sddmm: workspace too small");
    cudaStream_t st = (cudaStream_t)stream_;
    float* planes = (float*)workspace;
    k_sd_prep_weights<<<64, 256, 0, st>>>(W, K, Fi, Fo, planes);
    GNNML3_LAUNCH_CHECK();
    const int np = 2 * ((K * 32 + 127) / 128);
    CUtensorMap mW;
    int rc;
    if ((rc = make_map(&mW, planes, (int64_t)np * 128, 32, 32, 128))) return rc;
    CUtensorMap mX;
    const int win_rows = N < SD_WIN_ROWS ? (int)N : SD_WIN_ROWS;
    if ((rc = ts_make_xmap(&mX, X, N, Fi, ldx, win_rows))) return rc;
    SDParams P;
    P.tilewin = reinterpret_cast<const int2*>(win);
    P.win_rows = win_rows;
    P.rowptr = rowptr; P.col = col; P.X = X; P.ldx = ldx; P.Fi = Fi; P.GC = GC; P.ldg = ldg; P.Fo = Fo; P.K = K; P.N = N;
    P.n_tiles = cdiv(N, SD_ROWS);
    P.dea = dea;
    switch (K) {
        case 2: return sd_launch<2>(mW, mX, P, st);
        case 4: return sd_launch<4>(mW, mX, P, st);
        case 6: return sd_launch<6>(mW, mX, P, st);
        case 8: return sd_launch<8>(mW, mX, P, st);
        default: return set_err(GNNML3_ERR_INVALID, "fused_sddmm: no kernel for K=%d", K);
    }
}
