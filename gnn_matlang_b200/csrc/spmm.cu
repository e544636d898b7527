// K-channel segmented reduction (SpMM) and its edge-gradient counterpart (SDDMM) over the dst-sorted CSR.
//
// Reference semantics (libs/spect_conv.py:76-77,98-99 + PyG propagate, aggr='add'): for every support k
//     P_k(x)[t] = sum_{e: dst_e = t} edge_attr[e, k] * x[src_e]
// computed K times with a materialised [E, F] message tensor and an atomic scatter.  Here one pass over the
// row's edges reads each source row ONCE with 128-bit loads and applies all K channels into K register-held
// accumulators; there are no atomics and the summation order inside a row is the original edge order (the
// order of the reference's CPU scatter_add).  The kernel is HBM-bound: per row it streams the row's
// (K+1) * deg edge words, gathers deg source rows (L1/L2 hits: the sources of a graph's rows are its own
// <= ~100 contiguous nodes) and writes K*F outputs.
//
// Thread mapping: a row is owned by G lanes (G = power of two <= 32, chosen from F so that small feature
// widths do not idle a warp); each lane holds VEC consecutive features of CH chunks: f = (c*G + g)*VEC + v.
#include "common.cuh"

#include <cstdlib>

namespace gnnml3 {

template <int K, int WV>
__device__ __forceinline__ void load_weights(const float* __restrict__ p, float (&w)[K]) {
    static_assert(K % WV == 0, "weight vector width must divide the channel tile");
    if constexpr (WV == 4) {
#pragma unroll
        for (int k = 0; k < K; k += 4) {
            float4 v = ldg4(p + k);
            w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w;
        }
    } else if constexpr (WV == 2) {
#pragma unroll
        for (int k = 0; k < K; k += 2) {
            float2 v = ldg2(p + k);
            w[k] = v.x; w[k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) w[k] = __ldg(p + k);
    }
}

template <int VEC>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        float4 t = ldg4(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (VEC == 2) {
        float2 t = ldg2(p);
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = __ldg(p);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        st_na4(p, make_float4(v[0], v[1], v[2], v[3]));
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
        *p = v[0];
    }
}

template <int K, int WV, int VEC, int CH>
__global__ void __launch_bounds__(256)
k_spmm(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ eperm,
       const float* __restrict__ ea, const float* __restrict__ x, int64_t ldx, int N, int F, int G, int Kstride,
       float* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int rpw = 32 / G;
    const int64_t warp = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t row = warp * rpw + lane / G;
    const int g = lane & (G - 1);
    if (row >= N) return;

    float acc[K][CH][VEC];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[k][c][v] = 0.f;

    const int rs = __ldg(rowptr + row), re = __ldg(rowptr + row + 1);
    // U edges per iteration: all index, weight and source-row loads of the U edges are issued before the first FMA,
    // so a lane keeps U * (CH + K/WV) independent 64/128-bit loads in flight (rows are short -- 5..20 edges -- and
    // the index -> gather chain would otherwise serialise on memory latency).  Slots past the end of the row load a
    // clamped (valid) edge and skip the FMAs.
    constexpr int U = 2;   // measured on B200: 2 beats 4 (rows hold ~6 edges; slots past the row end still cost their loads)
    for (int p0 = rs; p0 < re; p0 += U) {
        int sidx[U], eidx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = min(p0 + u, re - 1);
            sidx[u] = __ldg(col + p);
            eidx[u] = eperm ? __ldg(eperm + p) : p;
        }
        float w[U][K], xv[U][CH][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            load_weights<K, WV>(ea + (int64_t)eidx[u] * Kstride, w[u]);
            const float* xr = x + (int64_t)sidx[u] * ldx;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int f0 = (c * G + g) * VEC;
                if (f0 < F) {
                    load_vec<VEC>(xr + f0, xv[u][c]);
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) xv[u][c][v] = 0.f;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (p0 + u < re) {       // keeps the reference's summation order: edges of a row in original order
#pragma unroll
                for (int c = 0; c < CH; ++c)
#pragma unroll
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) acc[k][c][v] = fmaf(w[u][k], xv[u][c][v], acc[k][c][v]);
            }
        }
    }
    float* orow = out + row * ldo;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int f0 = (c * G + g) * VEC;
        if (f0 < F) {
#pragma unroll
            for (int k = 0; k < K; ++k) store_vec<VEC>(orow + (int64_t)k * F + f0, acc[k][c]);
        }
    }
}

// SDDMM: dea[e(p), k] = <x[col[p]], g[t, k*F:(k+1)*F]>.  The row's K gradient slices stay in registers and
// each source row is read once per edge; the G lanes of a row reduce by xor-shuffles (all lanes of the warp
// run the same trip count so the full-mask shuffles are well defined).
template <int K, int VEC, int CH>
__global__ void __launch_bounds__(256)
k_sddmm(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ eperm,
        const float* __restrict__ x, int64_t ldx, const float* __restrict__ gin, int64_t ldg, int N, int F, int G,
        int Kstride, float* __restrict__ dea) {
    const int lane = threadIdx.x & 31;
    const int rpw = 32 / G;
    const int64_t warp = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t row = warp * rpw + lane / G;
    const int g = lane & (G - 1);
    const bool live = row < N;

    float gv[K][CH][VEC];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int f0 = (c * G + g) * VEC;
            if (live && f0 < F) {
                load_vec<VEC>(gin + row * ldg + (int64_t)k * F + f0, gv[k][c]);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) gv[k][c][v] = 0.f;
            }
        }
    int rs = 0, cnt = 0;
    if (live) {
        rs = __ldg(rowptr + row);
        cnt = __ldg(rowptr + row + 1) - rs;
    }
    int maxcnt = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, o));

    // where this lane's fully reduced values land after the reduce-scatter below
    constexpr int KP = K <= 1 ? 1 : (K <= 2 ? 2 : (K <= 4 ? 4 : 8));
    int kbase = 0, lowmask = 0, nkeep = KP;
    {
        int o = G >> 1, h = KP >> 1;
        while (h >= 1 && o >= 1) {
            if (g & o) kbase += h;
            nkeep = h;
            h >>= 1;
            o >>= 1;
        }
        if (o >= 1) lowmask = 2 * o - 1;
    }

    for (int i = 0; i < maxcnt; ++i) {
        const bool act = i < cnt;
        float part[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) part[k] = 0.f;
        int e = 0;
        if (act) {
            const int p = rs + i;
            const int s = __ldg(col + p);
            e = eperm ? __ldg(eperm + p) : p;
            const float* xr = x + (int64_t)s * ldx;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int f0 = (c * G + g) * VEC;
                if (f0 < F) {
                    float xv[VEC];
                    load_vec<VEC>(xr + f0, xv);
#pragma unroll
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) part[k] = fmaf(gv[k][c][v], xv[v], part[k]);
                }
            }
        }
        // reduce-scatter over the G lanes of the row: every step halves the number of live values per lane
        // (K/2 + K/4 + ... shuffles instead of K log2 G); once one value is left the remaining steps are plain adds
        {
            int o = G >> 1;
#pragma unroll
            for (int h = KP >> 1; h >= 1; h >>= 1) {
                if (o >= 1) {
                    const bool upper = (g & o) != 0;
#pragma unroll
                    for (int q = 0; q < h; ++q) {
                        const float keep = upper ? part[q + h] : part[q];
                        const float send = upper ? part[q] : part[q + h];
                        part[q] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                    }
                    o >>= 1;
                }
            }
            for (; o >= 1; o >>= 1) part[0] += __shfl_xor_sync(0xffffffffu, part[0], o);
        }
        if (act && (g & lowmask) == 0) {
            float* o = dea + (int64_t)e * Kstride + kbase;
#pragma unroll
            for (int q = 0; q < KP; ++q)
                if (q < nkeep && kbase + q < K) o[q] = part[q];
        }
    }
}

struct RowCfg {
    int vec, G, ch;
};

static const bool g_wide_rows = [] {
    const char* e = getenv("GNNML3_SPMM_WIDE");
    return e ? (e[0] != '0') : true;
}();

static bool pick_cfg(int F, bool aligned4, bool aligned2, RowCfg* c) {
    int vec = (F % 4 == 0 && aligned4) ? 4 : ((F % 2 == 0 && aligned2) ? 2 : 1);
    int units = F / vec;
    int G = 1;
    while (G < units && G < 32) G <<= 1;
    int ch = (units + G - 1) / G;
    if (ch == 3) ch = 4;
    if (ch > 4) return false;
    if (vec == 2 && ch > 1) {  // only (2,1) is instantiated: fall back to scalar lanes
        vec = 1;
        units = F;
        G = 32;
        ch = (units + 31) / 32;
        if (ch == 3) ch = 4;
        if (ch > 4) return false;
    }
    // fewer lanes per row, two chunks per lane: halves the per-edge index / address / loop overhead per FMA
    // (the kernels are instruction-issue and latency bound, not bandwidth bound) and doubles the rows per warp
    if (ch == 1 && G >= 8 && g_wide_rows) {
        G >>= 1;
        ch = 2;
    }
    c->vec = vec;
    c->G = G;
    c->ch = ch;
    return true;
}

}  // namespace gnnml3

using namespace gnnml3;

// (channel tile, weight-load width) pairs that are instantiated: WV divides the tile
#define DISPATCH_KW(KV, WVV, ...)                                                                         \
    if (KV == 1) { constexpr int K_ = 1, W_ = 1; __VA_ARGS__; }                                          \
    else if (KV == 2 && WVV == 2) { constexpr int K_ = 2, W_ = 2; __VA_ARGS__; }                         \
    else if (KV == 2) { constexpr int K_ = 2, W_ = 1; __VA_ARGS__; }                                     \
    else if (KV == 3) { constexpr int K_ = 3, W_ = 1; __VA_ARGS__; }                                     \
    else if (KV == 4 && WVV == 4) { constexpr int K_ = 4, W_ = 4; __VA_ARGS__; }                         \
    else if (KV == 4 && WVV == 2) { constexpr int K_ = 4, W_ = 2; __VA_ARGS__; }                         \
    else if (KV == 4) { constexpr int K_ = 4, W_ = 1; __VA_ARGS__; }                                     \
    else if (KV == 5) { constexpr int K_ = 5, W_ = 1; __VA_ARGS__; }                                     \
    else if (KV == 6 && WVV >= 2) { constexpr int K_ = 6, W_ = 2; __VA_ARGS__; }                         \
    else if (KV == 6) { constexpr int K_ = 6, W_ = 1; __VA_ARGS__; }                                     \
    else if (KV == 7) { constexpr int K_ = 7, W_ = 1; __VA_ARGS__; }                                     \
    else if (KV == 8 && WVV == 4) { constexpr int K_ = 8, W_ = 4; __VA_ARGS__; }                         \
    else if (KV == 8 && WVV == 2) { constexpr int K_ = 8, W_ = 2; __VA_ARGS__; }                         \
    else if (KV == 8) { constexpr int K_ = 8, W_ = 1; __VA_ARGS__; }                                     \
    else return set_err(GNNML3_ERR_INVALID, "internal: K tile %d", KV);

#define DISPATCH_K(KV, ...)                                                   \
    switch (KV) {                                                             \
        case 1: { constexpr int K_ = 1; __VA_ARGS__; } break;                 \
        case 2: { constexpr int K_ = 2; __VA_ARGS__; } break;                 \
        case 3: { constexpr int K_ = 3; __VA_ARGS__; } break;                 \
        case 4: { constexpr int K_ = 4; __VA_ARGS__; } break;                 \
        case 5: { constexpr int K_ = 5; __VA_ARGS__; } break;                 \
        case 6: { constexpr int K_ = 6; __VA_ARGS__; } break;                 \
        case 7: { constexpr int K_ = 7; __VA_ARGS__; } break;                 \
        case 8: { constexpr int K_ = 8; __VA_ARGS__; } break;                 \
        default: return set_err(GNNML3_ERR_INVALID, "internal: K tile %d", KV); \
    }

#define DISPATCH_VC(VEC, CH, ...)                                                          \
    if (VEC == 4 && CH == 1) { constexpr int V_ = 4, C_ = 1; __VA_ARGS__; }                \
    else if (VEC == 4 && CH == 2) { constexpr int V_ = 4, C_ = 2; __VA_ARGS__; }           \
    else if (VEC == 4 && CH == 4) { constexpr int V_ = 4, C_ = 4; __VA_ARGS__; }           \
    else if (VEC == 2 && CH == 1) { constexpr int V_ = 2, C_ = 1; __VA_ARGS__; }           \
    else if (VEC == 2 && CH == 2) { constexpr int V_ = 2, C_ = 2; __VA_ARGS__; }           \
    else if (VEC == 1 && CH == 1) { constexpr int V_ = 1, C_ = 1; __VA_ARGS__; }           \
    else if (VEC == 1 && CH == 2) { constexpr int V_ = 1, C_ = 2; __VA_ARGS__; }           \
    else if (VEC == 1 && CH == 4) { constexpr int V_ = 1, C_ = 4; __VA_ARGS__; }           \
    else return set_err(GNNML3_ERR_INVALID, "internal: no kernel for vec=%d ch=%d", VEC, CH);

// K is processed in register tiles of at most 8 channels (fewer for wide rows) so that the accumulators never
// spill; every tile re-walks the row (indices and source rows then hit L1/L2).  Even tiles are preferred so
// that the edge weights can be fetched with 64/128-bit loads.
static int k_tile_for(int K, const RowCfg& c) {
    const int per = c.vec * c.ch;  // accumulators per channel per lane
    int lim = 64 / per;
    if (lim > 8) lim = 8;
    if (lim < 1) lim = 1;
    if (K <= lim) return K;
    const int nt = (K + lim - 1) / lim;
    int kt = (K + nt - 1) / nt;
    if (K % 2 == 0 && kt % 2 == 1) kt = (kt + 1 <= lim) ? kt + 1 : kt - 1;
    return kt < 1 ? 1 : kt;
}

extern "C" int gnnml3_spmm_k(const int32_t* rowptr, const int32_t* col, const int32_t* eperm, const float* ea,
                             const float* x, int64_t ldx, int64_t N, int K, int F, float* out, int64_t ldo,
                             void* stream_) {
    GNNML3_REQUIRE(N > 0 && K > 0 && F > 0, "spmm_k: bad shape N=%lld K=%d F=%d", (long long)N, K, F);
    GNNML3_REQUIRE(rowptr && col && ea && x && out, "spmm_k: NULL pointer");
    GNNML3_REQUIRE(ldx >= F && ldo >= (int64_t)K * F, "spmm_k: leading dimensions too small");
    const bool a4 = ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0) && ldx % 4 == 0 && ldo % 4 == 0;
    const bool a2 = ((uintptr_t)x % 8 == 0) && ((uintptr_t)out % 8 == 0) && ldx % 2 == 0 && ldo % 2 == 0;
    RowCfg c;
    GNNML3_REQUIRE(pick_cfg(F, a4, a2, &c), "spmm_k: F=%d too wide for this build (max 512 aligned / 128 unaligned)", F);
    const int kt = k_tile_for(K, c);
    const int rpw = 32 / c.G;
    const int64_t warps = (N + rpw - 1) / rpw;
    const int blocks = (int)((warps + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream_;
    for (int k0 = 0; k0 < K;) {
        const int kk = (K - k0 < kt) ? (K - k0) : kt;
        // widest legal vector load of a tile's weights ea[e*K + k0 .. k0+kk)
        int wv = 1;
        if (kk % 4 == 0 && K % 4 == 0 && k0 % 4 == 0 && (uintptr_t)ea % 16 == 0) wv = 4;
        else if (kk % 2 == 0 && K % 2 == 0 && k0 % 2 == 0 && (uintptr_t)ea % 8 == 0) wv = 2;
        DISPATCH_KW(kk, wv, DISPATCH_VC(c.vec, c.ch,
            (k_spmm<K_, W_, V_, C_><<<blocks, 256, 0, st>>>(rowptr, col, eperm, ea + k0, x, ldx, (int)N, F, c.G, K,
                                                            out + (int64_t)k0 * F, ldo))));
        GNNML3_LAUNCH_CHECK();
        k0 += kk;
    }
    return GNNML3_OK;
}

extern "C" int gnnml3_sddmm_k(const int32_t* rowptr, const int32_t* col, const int32_t* eperm, const float* x,
                              int64_t ldx, const float* g, int64_t ldg, int64_t N, int K, int F, float* dea,
                              void* stream_) {
    GNNML3_REQUIRE(N > 0 && K > 0 && F > 0, "sddmm_k: bad shape");
    GNNML3_REQUIRE(rowptr && col && x && g && dea, "sddmm_k: NULL pointer");
    GNNML3_REQUIRE(ldx >= F && ldg >= (int64_t)K * F, "sddmm_k: leading dimensions too small");
    const bool a4 = ((uintptr_t)x % 16 == 0) && ((uintptr_t)g % 16 == 0) && ldx % 4 == 0 && ldg % 4 == 0;
    const bool a2 = ((uintptr_t)x % 8 == 0) && ((uintptr_t)g % 8 == 0) && ldx % 2 == 0 && ldg % 2 == 0;
    RowCfg c;
    GNNML3_REQUIRE(pick_cfg(F, a4, a2, &c), "sddmm_k: F=%d too wide for this build", F);
    const int kt = k_tile_for(K, c);
    const int rpw = 32 / c.G;
    const int64_t warps = (N + rpw - 1) / rpw;
    const int blocks = (int)((warps + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream_;
    for (int k0 = 0; k0 < K; k0 += kt) {
        const int kk = (K - k0 < kt) ? (K - k0) : kt;
        DISPATCH_K(kk, DISPATCH_VC(c.vec, c.ch,
            (k_sddmm<K_, V_, C_><<<blocks, 256, 0, st>>>(rowptr, col, eperm, x, ldx, g + (int64_t)k0 * F, ldg, (int)N,
                                                         F, c.G, K, dea + k0))));
        GNNML3_LAUNCH_CHECK();
    }
    return GNNML3_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// Project-first order of SpectConv (north_star: "or alternatively pre-projects and then aggregates, chosen by F_in/F_out"):
//   Y = x [W_0 .. W_{K-1}]  ([N, K*Fo], one tensor-core GEMM),  out[t, :] = sum_{p in row t} sum_k ea[e(p), k] * Y[col[p], k*Fo : (k+1)*Fo] (+ bias)
// Same mathematics as libs/spect_conv.py:70-80 with the sum over k moved inside the edge sum.  It gathers K*Fo floats per
// support entry instead of Fi, so it wins when the layer narrows (K*Fo*(E + 2N) < Fi*(E + 2NK) in HBM words).
// One warp per row, lanes over the output features (coalesced reads of the gathered Y rows), edge order, no atomics.
// ------------------------------------------------------------------------------------------------------------------------
namespace gnnml3 {

template <int FPL>      // output features per lane (Fo <= 32 * FPL)
__global__ void __launch_bounds__(256) k_spmm_projected(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ eperm,
                                                        const float* __restrict__ ea, int K, const float* __restrict__ Y, int64_t ldy, int N,
                                                        int Fo, const float* __restrict__ bias, float* __restrict__ out, int64_t ldo) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N) return;
    const int rs = __ldg(rowptr + warp), re = __ldg(rowptr + warp + 1);
    float acc[FPL];
#pragma unroll
    for (int i = 0; i < FPL; ++i) acc[i] = 0.f;
    for (int p = rs; p < re; ++p) {
        const int s = __ldg(col + p);
        const int e = eperm ? __ldg(eperm + p) : p;
        const float* yr = Y + (int64_t)s * ldy;
        const float* wr = ea + (int64_t)e * K;
        for (int k = 0; k < K; ++k) {
            const float w = __ldg(wr + k);
#pragma unroll
            for (int i = 0; i < FPL; ++i) {
                const int f = lane + 32 * i;
                if (f < Fo) acc[i] = fmaf(w, __ldg(yr + k * Fo + f), acc[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < FPL; ++i) {
        const int f = lane + 32 * i;
        if (f < Fo) out[(int64_t)warp * ldo + f] = acc[i] + (bias ? __ldg(bias + f) : 0.f);
    }
}

}  // namespace gnnml3

extern "C" int gnnml3_spmm_projected(const int32_t* rowptr, const int32_t* col, const int32_t* eperm, const float* ea, int K,
                                     const float* Y, int64_t ldy, int64_t N, int Fo, const float* bias, float* out, int64_t ldo,
                                     void* stream_) {
    GNNML3_REQUIRE(N > 0 && K > 0 && Fo > 0 && Fo <= 256, "spmm_projected: bad shape N=%lld K=%d Fo=%d (Fo <= 256)", (long long)N, K, Fo);
    GNNML3_REQUIRE(rowptr && col && ea && Y && out, "spmm_projected: NULL pointer");
    GNNML3_REQUIRE(ldy >= (int64_t)K * Fo && ldo >= Fo, "spmm_projected: leading dimensions too small");
    cudaStream_t st = (cudaStream_t)stream_;
    const int blocks = (int)((N * 32 + 255) / 256);
    const int fpl = (Fo + 31) / 32;
    if (fpl == 1) k_spmm_projected<1><<<blocks, 256, 0, st>>>(rowptr, col, eperm, ea, K, Y, ldy, (int)N, Fo, bias, out, ldo);
    else if (fpl == 2) k_spmm_projected<2><<<blocks, 256, 0, st>>>(rowptr, col, eperm, ea, K, Y, ldy, (int)N, Fo, bias, out, ldo);
    else if (fpl <= 4) k_spmm_projected<4><<<blocks, 256, 0, st>>>(rowptr, col, eperm, ea, K, Y, ldy, (int)N, Fo, bias, out, ldo);
    else k_spmm_projected<8><<<blocks, 256, 0, st>>>(rowptr, col, eperm, ea, K, Y, ldy, (int)N, Fo, bias, out, ldo);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}
