// Fused per-edge MLP of ML3Layer (reference libs/spect_conv.py:206-207):
//     ea' = relu( W4 [ relu(W1 ea) || tanh(W2 ea) * tanh(W3 ea) ] )         W1,W2,W3: [2K, K], W4: [K, 4K], no bias
// The reference runs 4 small GEMMs + 6 elementwise kernels and materialises [E, 4K] temporaries; here one
// thread owns one edge, the 10 K^2 weights sit in shared memory (broadcast reads) and nothing but ea and ea'
// touches HBM (2 * 4K bytes per edge).  The backward recomputes the activations per edge (no saved
// temporaries), and reduces the four weight gradients inside the block as register-tiled outer products over
// a shared-memory staging tile, then across blocks in a fixed order (deterministic, no atomics).
//
// Instantiated for even K <= 16 with nedgeoutput == nedgeinput (every configuration of the reference scripts);
// other shapes take the generic GEMM-composed path in gnn_matlang_b200/libs/spect_conv.py.
#include "edge_mlp.cuh"

#include <stdlib.h>

#include <atomic>

namespace gnnml3 {

template <int K>
__global__ void __launch_bounds__(256)
k_edge_mlp_fwd(const float* __restrict__ ea, const int* __restrict__ eperm, const float* __restrict__ w1,
               const float* __restrict__ w2, const float* __restrict__ w3, const float* __restrict__ w4, int64_t E,
               float* __restrict__ out) {
    using C = EMC<K>;
    __shared__ __align__(16) float sw123[C::W123];
    __shared__ __align__(16) float sw4[C::W4];
    load_weights_smem<K>(sw123, sw4, w1, w2, w3, w4);
    __syncthreads();
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t src = eperm ? (int64_t)__ldg(eperm + e) : e;
        float in[C::KP];
        load_edge_row<K>(ea + src * K, in);
        float tmp[C::T];
#pragma unroll
        for (int j = 0; j < C::H; ++j) {
            const float a1 = dot_row<K>(sw123 + (0 * C::H + j) * C::KP, in);
            const float a2 = dot_row<K>(sw123 + (1 * C::H + j) * C::KP, in);
            const float a3 = dot_row<K>(sw123 + (2 * C::H + j) * C::KP, in);
            tmp[j] = fmaxf(a1, 0.f);
            tmp[C::H + j] = tanh_fast(a2) * tanh_fast(a3);
        }
        float o[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < C::T; j += 4) {
                const float4 w = *reinterpret_cast<const float4*>(sw4 + k * C::T + j);
                a = fmaf(w.x, tmp[j], a);
                a = fmaf(w.y, tmp[j + 1], a);
                a = fmaf(w.z, tmp[j + 2], a);
                a = fmaf(w.w, tmp[j + 3], a);
            }
            o[k] = fmaxf(a, 0.f);
        }
        float* op = out + e * K;
#pragma unroll
        for (int k = 0; k < K; k += 2) *reinterpret_cast<float2*>(op + k) = make_float2(o[k], o[k + 1]);
    }
}

// Backward.  Phase 1 (thread = edge): recompute, back-propagate to the pre-activations, optionally emit d ea,
// and stage [d_pre4 | tmp | d_pre123 | in] in shared memory.  Phase 2 (thread = 4x4 tile of the weight-gradient
// matrices, GROUPS thread groups splitting the tile's edges): accumulate the outer products.
template <int K, bool DIN>
__global__ void __launch_bounds__(EMC<K>::THREADS)
k_edge_mlp_bwd(const float* __restrict__ ea, const int* __restrict__ eperm, const float* __restrict__ gout,
               const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ w3,
               const float* __restrict__ w4, int64_t E, float* __restrict__ dea, float* __restrict__ partial) {
    using C = EMC<K>;
    extern __shared__ __align__(16) float smem[];
    float* stage = smem;                                  // [TILE_E][S]
    float* sw123 = smem + (size_t)C::TILE_E * C::S;
    float* sw4 = sw123 + C::W123;
    load_weights_smem<K>(sw123, sw4, w1, w2, w3, w4);

    const int tid = threadIdx.x;
    const int tile = tid % C::NT, group = tid / C::NT;
    const bool accum = group < C::GROUPS;
    int offA, offB;
    if (tile < C::NT1) {
        offA = C::OFF_D4 + 4 * (tile / (C::T / 4));
        offB = C::OFF_TMP + 4 * (tile % (C::T / 4));
    } else {
        const int t2 = tile - C::NT1;
        offA = C::OFF_D + 4 * (t2 / (C::KP / 4));
        offB = C::OFF_IN + 4 * (t2 % (C::KP / 4));
    }
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    __syncthreads();

    const int64_t ntiles = (E + C::TILE_E - 1) / C::TILE_E;
    for (int64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        // ---------------- phase 1
        const int64_t e = tl * C::TILE_E + tid;
        float* row = stage + (size_t)tid * C::S;
        if (e < E) {
            const int64_t src = eperm ? (int64_t)__ldg(eperm + e) : e;
            float in[C::KP];
            load_edge_row<K>(ea + src * K, in);
            float tmp[C::T];
            uint32_t mask1 = 0;
#pragma unroll
            for (int j0 = 0; j0 < C::H; j0 += 4) {
                float t2v[4], t3v[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = j0 + jj;
                    const float a1 = dot_row<K>(sw123 + (0 * C::H + j) * C::KP, in);
                    const float a2 = dot_row<K>(sw123 + (1 * C::H + j) * C::KP, in);
                    const float a3 = dot_row<K>(sw123 + (2 * C::H + j) * C::KP, in);
                    tmp[j] = fmaxf(a1, 0.f);
                    if (a1 > 0.f) mask1 |= (1u << j);
                    t2v[jj] = tanh_fast(a2);
                    t3v[jj] = tanh_fast(a3);
                    tmp[C::H + j] = t2v[jj] * t3v[jj];
                }
                // park tanh values in the d_pre2 / d_pre3 slots until the upstream gradient is known
                *reinterpret_cast<float4*>(row + C::OFF_D + C::H + j0) = make_float4(t2v[0], t2v[1], t2v[2], t2v[3]);
                *reinterpret_cast<float4*>(row + C::OFF_D + 2 * C::H + j0) = make_float4(t3v[0], t3v[1], t3v[2], t3v[3]);
            }
#pragma unroll
            for (int j = 0; j < C::T; j += 4)
                *reinterpret_cast<float4*>(row + C::OFF_TMP + j) = make_float4(tmp[j], tmp[j + 1], tmp[j + 2], tmp[j + 3]);
#pragma unroll
            for (int i = 0; i < C::KP; i += 4)
                *reinterpret_cast<float4*>(row + C::OFF_IN + i) = make_float4(in[i], in[i + 1], in[i + 2], in[i + 3]);
            // d_pre4 = gout * relu'(pre4)
            float d4[C::KP];
            {
                float go[C::KP];
                load_edge_row<K>(gout + e * K, go);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float a = 0.f;
#pragma unroll
                    for (int j = 0; j < C::T; j += 4) {
                        const float4 w = *reinterpret_cast<const float4*>(sw4 + k * C::T + j);
                        a = fmaf(w.x, tmp[j], a);
                        a = fmaf(w.y, tmp[j + 1], a);
                        a = fmaf(w.z, tmp[j + 2], a);
                        a = fmaf(w.w, tmp[j + 3], a);
                    }
                    d4[k] = a > 0.f ? go[k] : 0.f;
                }
#pragma unroll
                for (int k = K; k < C::KP; ++k) d4[k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < C::KP; k += 4)
                *reinterpret_cast<float4*>(row + C::OFF_D4 + k) = make_float4(d4[k], d4[k + 1], d4[k + 2], d4[k + 3]);
            // d_tmp = W4^T d_pre4, then the three first-layer pre-activation gradients, 4 hidden units at a time
            float din[DIN ? C::KP : 1];
#pragma unroll
            for (int i = 0; i < (DIN ? C::KP : 1); ++i) din[i] = 0.f;
#pragma unroll
            for (int j0 = 0; j0 < C::H; j0 += 4) {
                float dt1[4] = {0.f, 0.f, 0.f, 0.f}, dt2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float4 wa = *reinterpret_cast<const float4*>(sw4 + k * C::T + j0);
                    const float4 wb = *reinterpret_cast<const float4*>(sw4 + k * C::T + C::H + j0);
                    dt1[0] = fmaf(d4[k], wa.x, dt1[0]); dt1[1] = fmaf(d4[k], wa.y, dt1[1]);
                    dt1[2] = fmaf(d4[k], wa.z, dt1[2]); dt1[3] = fmaf(d4[k], wa.w, dt1[3]);
                    dt2[0] = fmaf(d4[k], wb.x, dt2[0]); dt2[1] = fmaf(d4[k], wb.y, dt2[1]);
                    dt2[2] = fmaf(d4[k], wb.z, dt2[2]); dt2[3] = fmaf(d4[k], wb.w, dt2[3]);
                }
                const float4 t2 = *reinterpret_cast<const float4*>(row + C::OFF_D + C::H + j0);
                const float4 t3 = *reinterpret_cast<const float4*>(row + C::OFF_D + 2 * C::H + j0);
                const float t2a[4] = {t2.x, t2.y, t2.z, t2.w}, t3a[4] = {t3.x, t3.y, t3.z, t3.w};
                float d1[4], d2[4], d3[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    d1[jj] = ((mask1 >> (j0 + jj)) & 1u) ? dt1[jj] : 0.f;
                    d2[jj] = dt2[jj] * t3a[jj] * (1.f - t2a[jj] * t2a[jj]);
                    d3[jj] = dt2[jj] * t2a[jj] * (1.f - t3a[jj] * t3a[jj]);
                }
                *reinterpret_cast<float4*>(row + C::OFF_D + j0) = make_float4(d1[0], d1[1], d1[2], d1[3]);
                *reinterpret_cast<float4*>(row + C::OFF_D + C::H + j0) = make_float4(d2[0], d2[1], d2[2], d2[3]);
                *reinterpret_cast<float4*>(row + C::OFF_D + 2 * C::H + j0) = make_float4(d3[0], d3[1], d3[2], d3[3]);
                if constexpr (DIN) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float* r1 = sw123 + (0 * C::H + j0 + jj) * C::KP;
                        const float* r2 = sw123 + (1 * C::H + j0 + jj) * C::KP;
                        const float* r3 = sw123 + (2 * C::H + j0 + jj) * C::KP;
#pragma unroll
                        for (int i = 0; i < C::KP; i += 4) {
                            const float4 a = *reinterpret_cast<const float4*>(r1 + i);
                            const float4 b = *reinterpret_cast<const float4*>(r2 + i);
                            const float4 c = *reinterpret_cast<const float4*>(r3 + i);
                            din[i] += d1[jj] * a.x + d2[jj] * b.x + d3[jj] * c.x;
                            din[i + 1] += d1[jj] * a.y + d2[jj] * b.y + d3[jj] * c.y;
                            din[i + 2] += d1[jj] * a.z + d2[jj] * b.z + d3[jj] * c.z;
                            din[i + 3] += d1[jj] * a.w + d2[jj] * b.w + d3[jj] * c.w;
                        }
                    }
                }
            }
            if constexpr (C::DP > C::D) {
#pragma unroll
                for (int i = C::D; i < C::DP; ++i) row[C::OFF_D + i] = 0.f;
            }
            if constexpr (DIN) {
                float* dp = dea + src * K;
#pragma unroll
                for (int i = 0; i < K; i += 2) *reinterpret_cast<float2*>(dp + i) = make_float2(din[i], din[i + 1]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < C::ROW; i += 4) *reinterpret_cast<float4*>(row + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        // ---------------- phase 2
        if (accum) {
#pragma unroll 4
            for (int r = group; r < C::TILE_E; r += C::GROUPS) {
                const float4 a = *reinterpret_cast<const float4*>(stage + (size_t)r * C::S + offA);
                const float4 b = *reinterpret_cast<const float4*>(stage + (size_t)r * C::S + offB);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    // ---------------- reduce the groups in a fixed order and emit this block's partial (tile layout)
    float* red = stage;   // [GROUPS][NT*16]
    if (accum) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[(size_t)group * C::NT * 16 + tile * 16 + i * 4 + j] = acc[i][j];
    }
    __syncthreads();
    for (int i = tid; i < C::NT * 16; i += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < C::GROUPS; ++g) s += red[(size_t)g * C::NT * 16 + i];
        partial[(size_t)blockIdx.x * C::NT * 16 + i] = s;
    }
}

static int bwd_blocks(int64_t E, int tile_e) {
    int64_t t = (E + tile_e - 1) / tile_e;
    int64_t cap = kNumSMs;  // one resident block per SM (the staging tile is 50-200 KB)
    return (int)(t < 1 ? 1 : (t > cap ? cap : t));
}

// tensor-core generation (edge_mlp_tc.cu)
int edge_mlp_tc_fwd(const float* ea, const int32_t* eperm, const float* w1, const float* w2, const float* w3, const float* w4, int64_t E,
                    int K, float* out, cudaStream_t st);
size_t edge_mlp_tc_bwd_workspace_bytes(int K);
int edge_mlp_tc_bwd(const float* ea, const int32_t* eperm, const float* gout, const float* w1, const float* w2, const float* w3,
                    const float* w4, int64_t E, int K, float* dea, float* dw1, float* dw2, float* dw3, float* dw4, float* partial,
                    cudaStream_t st);

}  // namespace gnnml3

using namespace gnnml3;

// GNNML3_EDGE_TC=0 in the environment (or gnnml3_edge_mlp_set_tc(0)) selects the CUDA-core kernels of this file
static int g_edge_tc = [] {
    const char* e = getenv("GNNML3_EDGE_TC");
    return (e && e[0] == '0') ? 0 : 1;
}();
static std::atomic<long long> g_edge_paths[2];   // launches of [tensor-core, CUDA-core] edge-MLP kernels (forward + backward)

extern "C" int gnnml3_edge_mlp_set_tc(int enable) {
    const int old = g_edge_tc;
    if (enable == 0 || enable == 1) g_edge_tc = enable;
    return old;
}
extern "C" int gnnml3_edge_mlp_path_counts(long long* out2_host, int reset) {
    out2_host[0] = reset ? g_edge_paths[0].exchange(0) : g_edge_paths[0].load();
    out2_host[1] = reset ? g_edge_paths[1].exchange(0) : g_edge_paths[1].load();
    return GNNML3_OK;
}

#define DISPATCH_EVEN_K(KV, ...)                                   \
    switch (KV) {                                                  \
        case 2: { constexpr int K_ = 2; __VA_ARGS__; } break;      \
        case 4: { constexpr int K_ = 4; __VA_ARGS__; } break;      \
        case 6: { constexpr int K_ = 6; __VA_ARGS__; } break;      \
        case 8: { constexpr int K_ = 8; __VA_ARGS__; } break;      \
        case 10: { constexpr int K_ = 10; __VA_ARGS__; } break;    \
        case 12: { constexpr int K_ = 12; __VA_ARGS__; } break;    \
        case 14: { constexpr int K_ = 14; __VA_ARGS__; } break;    \
        case 16: { constexpr int K_ = 16; __VA_ARGS__; } break;    \
        default: return set_err(GNNML3_ERR_INVALID, "edge_mlp: K=%d not instantiated (even K <= 16)", KV); \
    }

extern "C" int gnnml3_edge_mlp_supported(int K, int Kout) { return (K == Kout && K >= 2 && K <= 16 && K % 2 == 0) ? 1 : 0; }

extern "C" int gnnml3_edge_mlp_fwd(const float* ea, const int32_t* eperm, const float* w1, const float* w2, const float* w3,
                                   const float* w4, int64_t E, int K, int Kout, float* out, void* stream_) {
    GNNML3_REQUIRE(gnnml3_edge_mlp_supported(K, Kout), "edge_mlp_fwd: unsupported K=%d Kout=%d (even K <= 16, Kout == K)", K, Kout);
    GNNML3_REQUIRE(E >= 0, "edge_mlp_fwd: E < 0");
    if (E == 0) return GNNML3_OK;
    GNNML3_REQUIRE(ea && w1 && w2 && w3 && w4 && out, "edge_mlp_fwd: NULL pointer");
    GNNML3_REQUIRE((uintptr_t)ea % 16 == 0 && (uintptr_t)out % 16 == 0, "edge_mlp_fwd: ea/out must be 16-byte aligned");
    if (g_edge_tc) {
        ++g_edge_paths[0];
        return edge_mlp_tc_fwd(ea, eperm, w1, w2, w3, w4, E, K, out, (cudaStream_t)stream_);
    }
    ++g_edge_paths[1];
    int64_t nb = (E + 255) / 256;
    if (nb > (int64_t)kNumSMs * 8) nb = (int64_t)kNumSMs * 8;
    DISPATCH_EVEN_K(K, (k_edge_mlp_fwd<K_><<<(int)nb, 256, 0, (cudaStream_t)stream_>>>(ea, eperm, w1, w2, w3, w4, E, out)));
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

extern "C" size_t gnnml3_edge_mlp_bwd_workspace_bytes(int64_t E, int K) {
    const int KP = pad4(K), T = 4 * K, DP = pad4(6 * K);
    const size_t nt = (size_t)(KP / 4) * (T / 4) + (size_t)(DP / 4) * (KP / 4);
    const size_t cc = (size_t)bwd_blocks(E, 256) * nt * 16 * sizeof(float);   // 256 = smallest tile = most blocks
    const size_t tc = edge_mlp_tc_bwd_workspace_bytes(K);                      // one partial per worker group
    return align_up(cc > tc ? cc : tc, 256);
}

template <int K>
static int launch_edge_bwd(const float* ea, const int32_t* eperm, const float* gout, const float* w1, const float* w2,
                           const float* w3, const float* w4, int64_t E, float* dea, float* dw1, float* dw2, float* dw3,
                           float* dw4, float* partial, cudaStream_t st) {
    using C = EMC<K>;
    static bool configured[64] = {};
    if (auto once_ = first_use_on_device(configured)) {
        GNNML3_CUDA(cudaFuncSetAttribute(k_edge_mlp_bwd<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::bwd_smem));
        GNNML3_CUDA(cudaFuncSetAttribute(k_edge_mlp_bwd<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::bwd_smem));
    }
    const int nb = bwd_blocks(E, C::TILE_E);
    if (dea)
        k_edge_mlp_bwd<K, true><<<nb, C::THREADS, C::bwd_smem, st>>>(ea, eperm, gout, w1, w2, w3, w4, E, dea, partial);
    else
        k_edge_mlp_bwd<K, false><<<nb, C::THREADS, C::bwd_smem, st>>>(ea, eperm, gout, w1, w2, w3, w4, E, dea, partial);
    GNNML3_LAUNCH_CHECK();
    k_edge_mlp_bwd_reduce<K><<<cdiv(C::NOUT, 32), 256, 0, st>>>(partial, nb, dw1, dw2, dw3, dw4);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

extern "C" int gnnml3_edge_mlp_bwd(const float* ea, const int32_t* eperm, const float* gout, const float* w1, const float* w2,
                                   const float* w3, const float* w4, int64_t E, int K, int Kout, float* dea, float* dw1,
                                   float* dw2, float* dw3, float* dw4, void* workspace, size_t workspace_bytes,
                                   void* stream_) {
    GNNML3_REQUIRE(gnnml3_edge_mlp_supported(K, Kout), "edge_mlp_bwd: unsupported K=%d Kout=%d", K, Kout);
    GNNML3_REQUIRE(E >= 0 && w1 && w2 && w3 && w4 && dw1 && dw2 && dw3 && dw4, "edge_mlp_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream_;
    if (E == 0) {
        GNNML3_CUDA(cudaMemsetAsync(dw1, 0, sizeof(float) * 2 * K * K, st));
        GNNML3_CUDA(cudaMemsetAsync(dw2, 0, sizeof(float) * 2 * K * K, st));
        GNNML3_CUDA(cudaMemsetAsync(dw3, 0, sizeof(float) * 2 * K * K, st));
        GNNML3_CUDA(cudaMemsetAsync(dw4, 0, sizeof(float) * 4 * K * K, st));
        return GNNML3_OK;
    }
    GNNML3_REQUIRE(ea && gout && workspace, "edge_mlp_bwd: NULL pointer");
    GNNML3_REQUIRE((uintptr_t)ea % 16 == 0 && (uintptr_t)gout % 16 == 0 && (dea == nullptr || (uintptr_t)dea % 16 == 0),
                   "edge_mlp_bwd: ea/gout/dea must be 16-byte aligned");
    if (workspace_bytes < gnnml3_edge_mlp_bwd_workspace_bytes(E, K))
        return set_err(GNNML3_ERR_WORKSPACE, "edge_mlp_bwd: workspace too small");
    if (g_edge_tc) {
        ++g_edge_paths[0];
        return edge_mlp_tc_bwd(ea, eperm, gout, w1, w2, w3, w4, E, K, dea, dw1, dw2, dw3, dw4, (float*)workspace, st);
    }
    ++g_edge_paths[1];
    DISPATCH_EVEN_K(K, return launch_edge_bwd<K_>(ea, eperm, gout, w1, w2, w3, w4, E, dea, dw1, dw2, dw3, dw4,
                                                   (float*)workspace, st));
    return GNNML3_OK;
}
