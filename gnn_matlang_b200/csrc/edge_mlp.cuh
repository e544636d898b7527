// Shapes and helpers shared by the two generations of the per-edge MLP kernels (edge_mlp.cu: CUDA-core version,
// edge_mlp_tc.cu: tcgen05 version).  Reference: libs/spect_conv.py:191-194,206-207.
#pragma once
#include "common.cuh"

namespace gnnml3 {

constexpr int pad4(int x) { return (x + 3) / 4 * 4; }

template <int K>
struct EMC {
    static constexpr int KP = pad4(K);
    static constexpr int H = 2 * K;        // hidden width of each of the three first-layer branches
    static constexpr int T = 4 * K;        // width of the concatenated activation
    static constexpr int D = 6 * K;        // d_pre1 | d_pre2 | d_pre3
    static constexpr int DP = pad4(D);
    // staged per-edge row of the backward: [d_pre4 : KP][tmp : T][d_pre123 : DP][in : KP]
    static constexpr int OFF_D4 = 0, OFF_TMP = KP, OFF_D = KP + T, OFF_IN = KP + T + DP, ROW = KP + T + DP + KP;
    static constexpr int S = ((ROW / 4) % 2 == 1) ? ROW : ROW + 4;   // S/4 odd: conflict-free 128-bit row stores
    static constexpr int NT1 = (KP / 4) * (T / 4);    // 4x4 tiles of M1 = d_pre4^T tmp      [KP x T]
    static constexpr int NT2 = (DP / 4) * (KP / 4);   // 4x4 tiles of M2 = d_pre123^T in     [DP x KP]
    static constexpr int NT = NT1 + NT2;
    static constexpr int THREADS = 256;   // backward block: 320 / 384 threads (fewer registers, more warps) spill and measured 11 % slower
    static constexpr int GROUPS = (THREADS / NT) < 1 ? 1 : (THREADS / NT);
    static constexpr int TILE_E = THREADS;
    static constexpr int W123 = 3 * H * KP;           // smem floats for W1..W3 (rows padded to KP)
    static constexpr int W4 = K * T;
    static constexpr int NOUT = K * T + 3 * H * K;    // number of weight-gradient entries
    static constexpr size_t bwd_smem = sizeof(float) * ((size_t)TILE_E * S + W123 + W4);
};

template <int K>
__device__ __forceinline__ void load_weights_smem(float* sw123, float* sw4, const float* __restrict__ w1,
                                                  const float* __restrict__ w2, const float* __restrict__ w3,
                                                  const float* __restrict__ w4) {
    using C = EMC<K>;
    for (int i = threadIdx.x; i < C::W123; i += blockDim.x) {
        const int m = i / (C::H * C::KP), r = (i / C::KP) % C::H, c = i % C::KP;
        const float* w = m == 0 ? w1 : (m == 1 ? w2 : w3);
        sw123[i] = (c < K) ? __ldg(w + r * K + c) : 0.f;
    }
    for (int i = threadIdx.x; i < C::W4; i += blockDim.x) sw4[i] = __ldg(w4 + i);
}

template <int K>
__device__ __forceinline__ void load_edge_row(const float* __restrict__ p, float (&in)[EMC<K>::KP]) {
    // K is even, rows are K floats: 8-byte aligned when the base is; 16-byte when K % 4 == 0
    if constexpr (K % 4 == 0) {
#pragma unroll
        for (int i = 0; i < K; i += 4) {
            float4 v = ldg4(p + i);
            in[i] = v.x; in[i + 1] = v.y; in[i + 2] = v.z; in[i + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < K; i += 2) {
            float2 v = ldg2(p + i);
            in[i] = v.x; in[i + 1] = v.y;
        }
#pragma unroll
        for (int i = K; i < EMC<K>::KP; ++i) in[i] = 0.f;
    }
}

template <int K>
__device__ __forceinline__ float dot_row(const float* __restrict__ wrow, const float (&in)[EMC<K>::KP]) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < EMC<K>::KP; i += 4) {
        const float4 w = *reinterpret_cast<const float4*>(wrow + i);
        a = fmaf(w.x, in[i], a);
        a = fmaf(w.y, in[i + 1], a);
        a = fmaf(w.z, in[i + 2], a);
        a = fmaf(w.w, in[i + 3], a);
    }
    return a;
}

// 32 outputs per block, 8 split-lanes per output summing every 8th block partial (independent loads), then a fixed tree.
template <int K>
__global__ void __launch_bounds__(256)
k_edge_mlp_bwd_reduce(const float* __restrict__ partial, int nblocks, float* __restrict__ dw1, float* __restrict__ dw2,
                      float* __restrict__ dw3, float* __restrict__ dw4) {
    using C = EMC<K>;
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx;
    int idx = 0;
    float* dst = nullptr;
    if (i < C::NOUT) {
        int r, c;
        if (i < K * C::T) {               // dW4[k][j] = M1[k][j]
            r = i / C::T;
            c = i % C::T;
            idx = ((r / 4) * (C::T / 4) + c / 4) * 16 + (r % 4) * 4 + (c % 4);
            dst = dw4 + i;
        } else {                          // dW{1,2,3}[j][i] = M2[m*2K + j][i]
            const int q = i - K * C::T;
            const int m = q / (C::H * K), jj = (q / K) % C::H;
            c = q % K;
            r = m * C::H + jj;
            idx = (C::NT1 + (r / 4) * (C::KP / 4) + c / 4) * 16 + (r % 4) * 4 + (c % 4);
            dst = (m == 0 ? dw1 : (m == 1 ? dw2 : dw3)) + jj * K + c;
        }
    }
    float s = 0.f;
    if (dst) {
        float s1 = 0.f, s2 = 0.f, s3 = 0.f;      // four loads in flight, fixed order
        int b = ty;
        for (; b + 24 < nblocks; b += 32) {
            s += __ldg(partial + (size_t)b * C::NT * 16 + idx);
            s1 += __ldg(partial + (size_t)(b + 8) * C::NT * 16 + idx);
            s2 += __ldg(partial + (size_t)(b + 16) * C::NT * 16 + idx);
            s3 += __ldg(partial + (size_t)(b + 24) * C::NT * 16 + idx);
        }
        for (; b < nblocks; b += 8) s += __ldg(partial + (size_t)b * C::NT * 16 + idx);
        s = (s + s1) + (s2 + s3);
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && dst) {
        float v = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) v += red[y][tx];
        *dst = v;
    }
}


}  // namespace gnnml3
