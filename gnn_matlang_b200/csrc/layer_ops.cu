// Node-side epilogue of ML3Layer and the graph readout.
//
//   ml3_act_fwd : y = [ relu(c) || tanh(p1) * tanh(p2) ]   from pre = [c | p1 | p2]   (reference libs/spect_conv.py:209-212:
//                 relu(conv1(...)), tanh(fc11 x) * tanh(fc12 x), torch.cat -- five elementwise launches + a cat there)
//   ml3_act_bwd : gradient of the above w.r.t. pre (recomputes the tanh values; nothing extra is saved)
//   segment_pool: global_add_pool / global_mean_pool (graph8c.py:277, Zinc12k.py:343, exp_classify.py:293,
//                 counting.py:370) over the contiguous node range of every graph of the batch -- a segmented
//                 sum in node order, no atomics.
#include "common.cuh"

#include <cstdlib>

namespace gnnml3 {

__global__ void k_ml3_act_fwd(const float* __restrict__ pre, int64_t ldp, int64_t N, int Fo, int G, float* __restrict__ y,
                              int64_t ldy) {
    const int W = Fo + G;
    const int64_t total = N * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / W;
        const int c = (int)(i - n * W);
        const float* p = pre + n * ldp;
        float v;
        if (c < Fo) {
            v = fmaxf(__ldg(p + c), 0.f);
        } else {
            v = tanhf(__ldg(p + c)) * tanhf(__ldg(p + c + G));
        }
        y[n * ldy + c] = v;
    }
}

// gpre[:, :Fo] = gy[:, :Fo] * (c > 0);  gpre[:, Fo+g] = gy[:, Fo+g] * t2 * (1 - t1^2);  gpre[:, Fo+G+g] = gy[:, Fo+g] * t1 * (1 - t2^2)
// gate_out (optional) receives a second copy of the 2G gate-gradient columns (row stride ldgate).
// The block also accumulates the column sums of gpre over its ACT_ROWS rows (the bias gradients) and writes them to
// colpart[block][Fo+2G]; a fixed-order second pass (k_colsum_finish) adds the blocks -- no extra pass over gpre.
constexpr int ACT_ROWS = 128;
__global__ void __launch_bounds__(256)
k_ml3_act_bwd(const float* __restrict__ pre, int64_t ldp, const float* __restrict__ gy, int64_t ldy, int64_t N,
              int Fo, int G, float* __restrict__ gpre, int64_t ldg, float* __restrict__ gate_out, int64_t ldgate,
              float* __restrict__ colpart) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int W = Fo + G, W2 = Fo + 2 * G;
    const int64_t r0 = (int64_t)blockIdx.x * ACT_ROWS;
    const int64_t r1 = min(N, r0 + ACT_ROWS);
    for (int c0 = 0; c0 < W; c0 += 32) {
        const int c = c0 + tx;
        float s1 = 0.f, s2 = 0.f;
        if (c < W) {
            for (int64_t n = r0 + ty; n < r1; n += 8) {
                const float* p = pre + n * ldp;
                const float g = __ldg(gy + n * ldy + c);
                if (c < Fo) {
                    const float v = __ldg(p + c) > 0.f ? g : 0.f;
                    gpre[n * ldg + c] = v;
                    s1 += v;
                } else {
                    const float t1 = tanhf(__ldg(p + c)), t2 = tanhf(__ldg(p + c + G));
                    const float g1 = g * t2 * (1.f - t1 * t1), g2 = g * t1 * (1.f - t2 * t2);
                    gpre[n * ldg + c] = g1;
                    gpre[n * ldg + c + G] = g2;
                    if (gate_out) {
                        gate_out[n * ldgate + (c - Fo)] = g1;
                        gate_out[n * ldgate + (c - Fo) + G] = g2;
                    }
                    s1 += g1;
                    s2 += g2;
                }
            }
        }
        if (colpart) {
            red[ty][tx] = s1;
            __syncthreads();
            if (ty == 0 && c < W) {
                float v = 0.f;
#pragma unroll
                for (int y = 0; y < 8; ++y) v += red[y][tx];
                colpart[(int64_t)blockIdx.x * W2 + c] = v;
            }
            __syncthreads();
            if (c >= Fo) {          // second gate column block
                red[ty][tx] = s2;
            }
            __syncthreads();
            if (ty == 0 && c >= Fo && c < W) {
                float v = 0.f;
#pragma unroll
                for (int y = 0; y < 8; ++y) v += red[y][tx];
                colpart[(int64_t)blockIdx.x * W2 + c + G] = v;
            }
            __syncthreads();
        }
    }
}

// Same gradient from the OUTPUTS of the fused layer kernel (fused_layer.cu): the ReLU mask is y > 0 and the gate factors
// t1 = tanh(p1), t2 = tanh(p2) were saved in aux [N, 2G], so `pre` is never materialised.  gpre uses the layout the fused
// dx kernel gathers from: conv gradient in columns [0, Fo), zero padding to Fo4 = ceil4(Fo), gate gradients
// [g1 | g2] in [Fo4, Fo4 + 2G), zero padding up to the row stride ldg (all 16-byte aligned blocks).  colpart as above,
// in the logical order [conv | g1 | g2].
constexpr int ACTY_ROWS = 128;      // rows per block; one thread per (row, 4-column group): 128-bit loads and stores
constexpr int ACTY_MAXLD = 72;      // widest gpre row staged in shared memory for the column sums
__global__ void __launch_bounds__(256)
k_ml3_act_bwd_y(const float* __restrict__ y, int64_t ldy, const float* __restrict__ aux, int64_t ldaux,
                const float* __restrict__ gy, int64_t ldgy, int64_t N, int Fo, int G, float* __restrict__ gpre, int64_t ldg,
                float* __restrict__ colpart, int vec_in) {
    __shared__ __align__(16) float tile[ACTY_ROWS * ACTY_MAXLD];
    const int Fo4 = (Fo + 3) / 4 * 4, W2 = Fo + 2 * G;
    const int ngroups = (int)(ldg / 4);
    const int64_t r0 = (int64_t)blockIdx.x * ACTY_ROWS;
    const int rows = (int)min((int64_t)ACTY_ROWS, N - r0);
    for (int i = threadIdx.x; i < rows * ngroups; i += blockDim.x) {
        const int rl = i / ngroups, cg = i - rl * ngroups, c = 4 * cg;
        const int64_t n = r0 + rl;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (c < Fo && vec_in && c + 3 < (int)ldy && c + 3 < (int)ldgy) {
            // (the group that straddles Fo reads the whole 16-byte chunk too -- it lies inside the row -- and masks the tail:
            // its scalar path made every warp execute both branches)
            const float4 yy = ldg4(y + n * ldy + c), gg = ldg4(gy + n * ldgy + c);
            v[0] = yy.x > 0.f ? gg.x : 0.f;
            v[1] = (c + 1 < Fo && yy.y > 0.f) ? gg.y : 0.f;
            v[2] = (c + 2 < Fo && yy.z > 0.f) ? gg.z : 0.f;
            v[3] = (c + 3 < Fo && yy.w > 0.f) ? gg.w : 0.f;
        } else if (c >= Fo4 && G == 2 && c == Fo4 && vec_in && ldaux % 4 == 0 && ((uintptr_t)aux & 15) == 0 && Fo + 1 < (int)ldgy) {
            // the common gate block (G = 2): [g1_0 g1_1 g2_0 g2_1] from one 16-byte load of the saved tanh factors
            const float4 t = ldg4(aux + n * ldaux);                    // t1_0 t1_1 t2_0 t2_1
            const float g0 = __ldg(gy + n * ldgy + Fo), g1 = __ldg(gy + n * ldgy + Fo + 1);
            v[0] = g0 * t.z * (1.f - t.x * t.x);
            v[1] = g1 * t.w * (1.f - t.y * t.y);
            v[2] = g0 * t.x * (1.f - t.z * t.z);
            v[3] = g1 * t.y * (1.f - t.w * t.w);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int cc = c + k;
                if (cc < Fo) {
                    v[k] = __ldg(y + n * ldy + cc) > 0.f ? __ldg(gy + n * ldgy + cc) : 0.f;
                } else if (cc >= Fo4 && cc < Fo4 + 2 * G) {
                    const int j = (cc - Fo4) % G, second = (cc - Fo4) / G;
                    const float g = __ldg(gy + n * ldgy + Fo + j);
                    const float t1 = __ldg(aux + n * ldaux + j), t2 = __ldg(aux + n * ldaux + G + j);
                    v[k] = second ? g * t1 * (1.f - t2 * t2) : g * t2 * (1.f - t1 * t1);
                }
            }
        }
        const float4 o = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(gpre + n * ldg + c) = o;
        if (colpart) *reinterpret_cast<float4*>(tile + rl * (int)ldg + c) = o;
    }
    if (colpart) {
        // column sums of the staged tile in two fixed-order levels: (segment of rows, column) per thread, then the segments of a
        // column (one thread per column walking all 128 rows serially was half of the kernel's time)
        __shared__ float seg_sum[256];
        __syncthreads();
        const int ld = (int)ldg;
        const int nseg = (int)blockDim.x / ld;                       // ld <= ACTY_MAXLD = 72: at least 3 segments
        const int rps = (ACTY_ROWS + nseg - 1) / nseg;
        const int sg = threadIdx.x / ld, c = threadIdx.x - sg * ld;
        if (sg < nseg) {
            float t = 0.f;
            const int r1 = min(rows, (sg + 1) * rps);
            for (int rl = sg * rps; rl < r1; ++rl) t += tile[rl * ld + c];
            seg_sum[sg * ld + c] = t;
        }
        __syncthreads();
        if ((int)threadIdx.x < ld) {
            const int cc = threadIdx.x;
            int lc = -1;
            if (cc < Fo) lc = cc;
            else if (cc >= Fo4 && cc < Fo4 + 2 * G) lc = Fo + (cc - Fo4);
            if (lc >= 0) {
                float t = 0.f;
                for (int g2 = 0; g2 < nseg; ++g2) t += seg_sum[g2 * ld + cc];      // fixed order: deterministic
                colpart[(int64_t)blockIdx.x * W2 + lc] = t;
            }
        }
    }
}

// Streaming form of k_ml3_act_bwd_y for the aligned shapes of the GNNML3 layers (Fo4 = 4 GCV floats with GCV a power of two,
// gate width 2 or a multiple of 4): a thread keeps ONE 4-column group and walks the block's rows with a fixed stride, so
//   * all of its 128-bit loads are independent and issued before the first use (the general kernel recomputes (row, group)
//     by a division per element and kept ~4 loads in flight: 2.3 TB/s),
//   * the column sums accumulate in registers (no shared-memory tile, no second pass over it): shuffle tree over the lanes
//     that share the group, then the eight warps in fixed order -- deterministic.
// Same values in gpre bit for bit (same expressions); same per-block partial layout, finished by k_colsum_finish.
// GW (gate width 2 only): the kernel also contracts the gate gradients with the layer input -- dW11 / dW12 = x^T [g1 | g2], 4 columns
// against Fi <= 32 -- while the rows are in flight: x adds 128 bytes per row to the stream, the 16 x 4 products per row are free
// next to the loads, and the separate narrow contraction (k_gemm_tn_narrow + its partial reduction: 25 us per layer on the ZINC
// step) disappears.  Partials: 128 more columns per block (x column i, gate column g at W2 + 4 i + g), finished by the same
// launch as the bias sums (k_colsum_finish_gw).
constexpr int ACTY_GW_COLS = 128;
template <int GCV, bool GW>
__global__ void __launch_bounds__(256)
k_ml3_act_bwd_y_v(const float* __restrict__ y, int64_t ldy, const float* __restrict__ aux, int64_t ldaux,
                  const float* __restrict__ gy, int64_t ldgy, int64_t N, int Fo, int G, float* __restrict__ gpre, int64_t ldg,
                  float* __restrict__ colpart, int pstride, const float* __restrict__ x, int64_t ldx, int Fi) {
    constexpr int RPP = 256 / GCV;                  // rows per pass of the conv block
    constexpr int NPASS = ACTY_ROWS / RPP;          // = GCV / 2
    __shared__ float red[8][GW ? 132 : 68];
    __shared__ float4 sgate[GW ? ACTY_ROWS : 1];
    const int Fo4 = 4 * GCV, W2 = Fo + 2 * G;
    const int64_t r0 = (int64_t)blockIdx.x * ACTY_ROWS;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    // ---- conv block: d pre = (y > 0) ? gy : 0
    {
        const int cg = t % GCV, rl0 = t / GCV, c = 4 * cg;
        float4 yy[NPASS], gg[NPASS];
#pragma unroll
        for (int j = 0; j < NPASS; ++j) {
            const int64_t n = r0 + rl0 + j * RPP;
            if (n < N) {
                yy[j] = ldg4(y + n * ldy + c);
                gg[j] = ldg4(gy + n * ldgy + c);
            }
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NPASS; ++j) {
            const int64_t n = r0 + rl0 + j * RPP;
            if (n < N) {
                float4 o;
                o.x = yy[j].x > 0.f ? gg[j].x : 0.f;
                o.y = (c + 1 < Fo && yy[j].y > 0.f) ? gg[j].y : 0.f;
                o.z = (c + 2 < Fo && yy[j].z > 0.f) ? gg[j].z : 0.f;
                o.w = (c + 3 < Fo && yy[j].w > 0.f) ? gg[j].w : 0.f;
                if (c >= Fo) o.x = 0.f;
                *reinterpret_cast<float4*>(gpre + n * ldg + c) = o;
                acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            }
        }
        if (colpart) {
#pragma unroll
            for (int o = GCV; o < 32; o <<= 1) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
                acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
            }
            if (GCV >= 32 || lane < GCV) *reinterpret_cast<float4*>(&red[warp][4 * (lane % GCV)]) = acc;
            __syncthreads();
            if (t < Fo) {
                float v = 0.f;
                if (GCV < 32) {
#pragma unroll
                    for (int w = 0; w < 8; ++w) v += red[w][t];
                }
                colpart[(int64_t)blockIdx.x * pstride + t] = v;
            }
            __syncthreads();
        }
    }
    // ---- gate block: [g1 | g2] at columns Fo4 .. Fo4 + 2G
    if (G == 2) {
        float4 xr[4];
        if constexpr (GW) {                          // this thread's 4-column group of the layer input, rows rl + 32 j
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t nx = r0 + (t >> 3) + 32 * j;
                xr[j] = (nx < N && 4 * (t & 7) < Fi) ? ldg4(x + nx * ldx + 4 * (t & 7)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t n = r0 + t;
        if (t < ACTY_ROWS && n < N) {
            const float4 tt = ldg4(aux + n * ldaux);                    // t1_0 t1_1 t2_0 t2_1
            const float g0 = __ldg(gy + n * ldgy + Fo), g1 = __ldg(gy + n * ldgy + Fo + 1);
            o.x = g0 * tt.z * (1.f - tt.x * tt.x);
            o.y = g1 * tt.w * (1.f - tt.y * tt.y);
            o.z = g0 * tt.x * (1.f - tt.z * tt.z);
            o.w = g1 * tt.y * (1.f - tt.w * tt.w);
            *reinterpret_cast<float4*>(gpre + n * ldg + Fo4) = o;
        }
        if constexpr (GW) {
            if (t < ACTY_ROWS) sgate[t] = o;         // zeros for rows beyond N
        }
        if (colpart) {
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                o.x += __shfl_xor_sync(0xffffffffu, o.x, s);
                o.y += __shfl_xor_sync(0xffffffffu, o.y, s);
                o.z += __shfl_xor_sync(0xffffffffu, o.z, s);
                o.w += __shfl_xor_sync(0xffffffffu, o.w, s);
            }
            if (lane == 0 && warp < 4) *reinterpret_cast<float4*>(&red[warp][0]) = o;
            __syncthreads();
            if (t < 4) colpart[(int64_t)blockIdx.x * pstride + Fo + t] = ((red[0][t] + red[1][t]) + red[2][t]) + red[3][t];
        }
        if constexpr (GW) {
            __syncthreads();                         // sgate complete; red free again
            float a[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int g = 0; g < 4; ++g) a[i][g] = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 gq = sgate[(t >> 3) + 32 * j];
                const float xv[4] = {xr[j].x, xr[j].y, xr[j].z, xr[j].w}, gv[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int g = 0; g < 4; ++g) a[i][g] = fmaf(xv[i], gv[g], a[i][g]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    a[i][g] += __shfl_xor_sync(0xffffffffu, a[i][g], 8);
                    a[i][g] += __shfl_xor_sync(0xffffffffu, a[i][g], 16);
                }
            if (lane < 8) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(&red[warp][16 * lane + 4 * i]) = make_float4(a[i][0], a[i][1], a[i][2], a[i][3]);
            }
            __syncthreads();
            if (t < ACTY_GW_COLS) {                  // column t = 4 * (x column) + gate column, fixed order over the warps
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) v += red[w][t];
                colpart[(int64_t)blockIdx.x * pstride + W2 + t] = v;
            }
        }
    } else if (G > 0) {
        // G % 4 == 0, Fo % 4 == 0: GG = G / 2 groups per row (a power of two <= 16), the first G / 4 belong to g1
        const int GG = G >> 1, rpp = 256 / GG, npass = ACTY_ROWS / rpp;
        const int gi = t % GG, rl0 = t / GG;
        const int second = gi >= (G >> 2) ? 1 : 0, j0 = 4 * (gi - second * (G >> 2));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < npass; ++j) {
            const int64_t n = r0 + rl0 + j * rpp;
            if (n < N) {
                const float4 g = ldg4(gy + n * ldgy + Fo + j0), t1 = ldg4(aux + n * ldaux + j0), t2 = ldg4(aux + n * ldaux + G + j0);
                float4 o;
                if (second) {
                    o.x = g.x * t1.x * (1.f - t2.x * t2.x); o.y = g.y * t1.y * (1.f - t2.y * t2.y);
                    o.z = g.z * t1.z * (1.f - t2.z * t2.z); o.w = g.w * t1.w * (1.f - t2.w * t2.w);
                } else {
                    o.x = g.x * t2.x * (1.f - t1.x * t1.x); o.y = g.y * t2.y * (1.f - t1.y * t1.y);
                    o.z = g.z * t2.z * (1.f - t1.z * t1.z); o.w = g.w * t2.w * (1.f - t1.w * t1.w);
                }
                *reinterpret_cast<float4*>(gpre + n * ldg + Fo4 + 4 * gi) = o;
                acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            }
        }
        if (colpart) {
            for (int s = GG; s < 32; s <<= 1) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
                acc.z += __shfl_xor_sync(0xffffffffu, acc.z, s);
                acc.w += __shfl_xor_sync(0xffffffffu, acc.w, s);
            }
            if (lane < GG) *reinterpret_cast<float4*>(&red[warp][4 * lane]) = acc;
            __syncthreads();
            if (t < 2 * G) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) v += red[w][t];
                colpart[(int64_t)blockIdx.x * pstride + Fo + t] = v;
            }
        }
    }
}

// out[c] = sum_b part[b][c]: one warp per column, lanes stride over the blocks, fixed-order shuffle tree (deterministic)
__global__ void k_colsum_finish(const float* __restrict__ part, int nblocks, int W2, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= W2) return;
    float s = 0.f;
    for (int b = lane; b < nblocks; b += 32) s += part[(int64_t)b * W2 + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[c] = s;
}

// bias sums and gate weight gradients in one finish launch: column c < W2 -> out[c]; column W2 + 4 i + g -> dW11[g][i] (g < 2) or
// dW12[g - 2][i], x columns i < Fi only
__global__ void __launch_bounds__(128)
k_colsum_finish_gw(const float* __restrict__ part, int nblocks, int pstride, int W2, int Fi, float* __restrict__ out,
                   float* __restrict__ dw11, float* __restrict__ dw12) {
    // one BLOCK per column (one warp per column walked ~46 dependent-free but serial loads per lane: 10 us per launch)
    __shared__ float red[4];
    const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = c - W2, i = j >> 2, g = j & 3;
    if (c >= W2 && i >= Fi) return;
    float s = 0.f;
    for (int b = threadIdx.x; b < nblocks; b += 128) s += part[(int64_t)b * pstride + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        s = (red[0] + red[1]) + (red[2] + red[3]);
        if (c < W2) { if (out) out[c] = s; }
        else if (g < 2) dw11[g * Fi + i] = s;
        else dw12[(g - 2) * Fi + i] = s;
    }
}

// one warp per (graph, 32-feature chunk)
__global__ void __launch_bounds__(256) k_segment_pool_fwd(const float* __restrict__ x, int64_t ldx, const int* __restrict__ gptr,
                                                          int B, int F, int mean, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int chunks = (F + 31) / 32;
    const int64_t w = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (int64_t)B * chunks) return;
    const int b = (int)(w / chunks), f = (int)(w % chunks) * 32 + lane;
    const int n0 = __ldg(gptr + b), n1 = __ldg(gptr + b + 1);
    if (f >= F) return;
    float s = 0.f;
    for (int n = n0; n < n1; ++n) s += __ldg(x + (int64_t)n * ldx + f);
    if (mean) s /= (float)max(n1 - n0, 1);
    out[(int64_t)b * F + f] = s;
}

__global__ void __launch_bounds__(256) k_segment_pool_bwd(const float* __restrict__ gout, const int* __restrict__ gptr, int B,
                                                          int F, int mean, float* __restrict__ gx, int64_t ldx) {
    const int lane = threadIdx.x & 31;
    const int chunks = (F + 31) / 32;
    const int64_t w = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (int64_t)B * chunks) return;
    const int b = (int)(w / chunks), f = (int)(w % chunks) * 32 + lane;
    const int n0 = __ldg(gptr + b), n1 = __ldg(gptr + b + 1);
    if (f >= F) return;
    float g = __ldg(gout + (int64_t)b * F + f);
    if (mean) g /= (float)max(n1 - n0, 1);
    for (int n = n0; n < n1; ++n) gx[(int64_t)n * ldx + f] = g;
}

// PyG global_max_pool (enzymes.py:340,384: read-out of the GNNML3 variants): per-graph maximum over the node range and the
// node that attains it (first one on ties, as a sequential scan would); an empty graph gives 0 / -1 (torch_scatter's fill).
__global__ void __launch_bounds__(256) k_segment_max_fwd(const float* __restrict__ x, int64_t ldx, const int* __restrict__ gptr,
                                                         int B, int F, float* __restrict__ out, int* __restrict__ arg) {
    const int lane = threadIdx.x & 31;
    const int chunks = (F + 31) / 32;
    const int64_t w = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (int64_t)B * chunks) return;
    const int b = (int)(w / chunks), f = (int)(w % chunks) * 32 + lane;
    const int n0 = __ldg(gptr + b), n1 = __ldg(gptr + b + 1);
    if (f >= F) return;
    float m = 0.f;
    int a = -1;
    for (int n = n0; n < n1; ++n) {
        const float v = __ldg(x + (int64_t)n * ldx + f);
        if (a < 0 || v > m) {
            m = v;
            a = n;
        }
    }
    out[(int64_t)b * F + f] = m;
    arg[(int64_t)b * F + f] = a;
}

// gx[n, f] = gout[b, f] if n == arg[b, f] else 0   (every element of gx written exactly once: no atomics, no pre-zeroing)
__global__ void __launch_bounds__(256) k_segment_max_bwd(const float* __restrict__ gout, const int* __restrict__ arg,
                                                         const int* __restrict__ gptr, int B, int F, float* __restrict__ gx, int64_t ldx) {
    const int lane = threadIdx.x & 31;
    const int chunks = (F + 31) / 32;
    const int64_t w = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (int64_t)B * chunks) return;
    const int b = (int)(w / chunks), f = (int)(w % chunks) * 32 + lane;
    const int n0 = __ldg(gptr + b), n1 = __ldg(gptr + b + 1);
    if (f >= F) return;
    const float g = __ldg(gout + (int64_t)b * F + f);
    const int a = __ldg(arg + (int64_t)b * F + f);
    for (int n = n0; n < n1; ++n) gx[(int64_t)n * ldx + f] = n == a ? g : 0.f;
}

static int ew_grid(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace gnnml3

using namespace gnnml3;

extern "C" int gnnml3_ml3_act_fwd(const float* pre, int64_t ldp, int64_t N, int Fo, int G, float* y, int64_t ldy, void* stream_) {
    GNNML3_REQUIRE(N >= 0 && Fo >= 0 && G >= 0 && Fo + G > 0, "ml3_act_fwd: bad shape");
    if (N == 0) return GNNML3_OK;
    GNNML3_REQUIRE(pre && y && ldp >= Fo + 2 * G && ldy >= Fo + G, "ml3_act_fwd: bad arguments");
    k_ml3_act_fwd<<<ew_grid(N * (Fo + G)), 256, 0, (cudaStream_t)stream_>>>(pre, ldp, N, Fo, G, y, ldy);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

extern "C" size_t gnnml3_ml3_act_bwd_workspace_bytes(int64_t N, int Fo, int G) {
    return align_up((size_t)cdiv(N > 0 ? N : 1, ACT_ROWS) * (Fo + 2 * G + 128 /* gate weight-gradient partials */) * sizeof(float), 256);
}

extern "C" int gnnml3_ml3_act_bwd(const float* pre, int64_t ldp, const float* gy, int64_t ldy, int64_t N, int Fo, int G,
                                  float* gpre, int64_t ldg, float* gate_out, int64_t ldgate, float* colsum,
                                  void* workspace, size_t workspace_bytes, void* stream_) {
    GNNML3_REQUIRE(N >= 0 && Fo >= 0 && G >= 0 && Fo + G > 0, "ml3_act_bwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream_;
    if (N == 0) {
        if (colsum) GNNML3_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * (Fo + 2 * G), st));
        return GNNML3_OK;
    }
    GNNML3_REQUIRE(pre && gy && gpre && ldp >= Fo + 2 * G && ldy >= Fo + G && ldg >= Fo + 2 * G, "ml3_act_bwd: bad arguments");
    GNNML3_REQUIRE(gate_out == nullptr || ldgate >= 2 * G, "ml3_act_bwd: ldgate too small");
    float* part = nullptr;
    if (colsum) {
        GNNML3_REQUIRE(workspace, "ml3_act_bwd: workspace required for the column sums");
        if (workspace_bytes < gnnml3_ml3_act_bwd_workspace_bytes(N, Fo, G))
            return set_err(GNNML3_ERR_WORKSPACE, "ml3_act_bwd: workspace too small");
        part = (float*)workspace;
    }
    const int nb = cdiv(N, ACT_ROWS);
    k_ml3_act_bwd<<<nb, 256, 0, st>>>(pre, ldp, gy, ldy, N, Fo, G, gpre, ldg, gate_out, ldgate, part);
    GNNML3_LAUNCH_CHECK();
    if (colsum) {
        k_colsum_finish<<<cdiv(Fo + 2 * G, 8), 256, 0, st>>>(part, nb, Fo + 2 * G, colsum);
        GNNML3_LAUNCH_CHECK();
    }
    return GNNML3_OK;
}

extern "C" int gnnml3_segment_pool_fwd(const float* x, int64_t ldx, const int32_t* graph_ptr, int B, int F, int mean,
                                       float* out, void* stream_) {
    GNNML3_REQUIRE(B >= 0 && F > 0, "segment_pool_fwd: bad shape");
    if (B == 0) return GNNML3_OK;
    GNNML3_REQUIRE(x && graph_ptr && out && ldx >= F, "segment_pool_fwd: bad arguments");
    const int64_t warps = (int64_t)B * ((F + 31) / 32);
    k_segment_pool_fwd<<<(int)((warps + 7) / 8), 256, 0, (cudaStream_t)stream_>>>(x, ldx, graph_ptr, B, F, mean, out);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

extern "C" int gnnml3_segment_max_fwd(const float* x, int64_t ldx, const int32_t* graph_ptr, int B, int F, float* out, int32_t* arg,
                                      void* stream_) {
    GNNML3_REQUIRE(B >= 0 && F > 0, "segment_max_fwd: bad shape");
    if (B == 0) return GNNML3_OK;
    GNNML3_REQUIRE(x && graph_ptr && out && arg && ldx >= F, "segment_max_fwd: bad arguments");
    const int64_t warps = (int64_t)B * ((F + 31) / 32);
    k_segment_max_fwd<<<(int)((warps + 7) / 8), 256, 0, (cudaStream_t)stream_>>>(x, ldx, graph_ptr, B, F, out, arg);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

extern "C" int gnnml3_segment_max_bwd(const float* gout, const int32_t* arg, const int32_t* graph_ptr, int B, int F, float* gx,
                                      int64_t ldx, void* stream_) {
    GNNML3_REQUIRE(B >= 0 && F > 0, "segment_max_bwd: bad shape");
    if (B == 0) return GNNML3_OK;
    GNNML3_REQUIRE(gout && arg && graph_ptr && gx && ldx >= F, "segment_max_bwd: bad arguments");
    const int64_t warps = (int64_t)B * ((F + 31) / 32);
    k_segment_max_bwd<<<(int)((warps + 7) / 8), 256, 0, (cudaStream_t)stream_>>>(gout, arg, graph_ptr, B, F, gx, ldx);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

extern "C" int gnnml3_segment_pool_bwd(const float* gout, const int32_t* graph_ptr, int B, int F, int mean, float* gx,
                                       int64_t ldx, void* stream_) {
    GNNML3_REQUIRE(B >= 0 && F > 0, "segment_pool_bwd: bad shape");
    if (B == 0) return GNNML3_OK;
    GNNML3_REQUIRE(gout && graph_ptr && gx && ldx >= F, "segment_pool_bwd: bad arguments");
    const int64_t warps = (int64_t)B * ((F + 31) / 32);
    k_segment_pool_bwd<<<(int)((warps + 7) / 8), 256, 0, (cudaStream_t)stream_>>>(gout, graph_ptr, B, F, mean, gx, ldx);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

namespace gnnml3 {
static int act_bwd_y_impl(const float* y, int64_t ldy, const float* aux, int64_t ldaux, const float* gy, int64_t ldgy, int64_t N, int Fo,
                          int G, float* gpre, int64_t ldg, float* colsum, const float* x, int64_t ldx, int Fi, float* dw11, float* dw12,
                          int* gate_dw_done, void* workspace, size_t workspace_bytes, void* stream_);
// layer_api.cu: d pre + bias sums and, when the shape allows (gate width 2, Fi <= 32, aligned x, N > 0), the gate weight gradients
// dW11 / dW12 [G, Fi] in the same two launches; *gate_dw_done tells the caller whether they were written
int ml3_act_bwd_y_gw(const float* y, int64_t ldy, const float* aux, int64_t ldaux, const float* gy, int64_t ldgy, int64_t N, int Fo, int G,
                     float* gpre, int64_t ldg, float* colsum, const float* x, int64_t ldx, int Fi, float* dw11, float* dw12,
                     int* gate_dw_done, void* workspace, size_t workspace_bytes, void* stream) {
    return act_bwd_y_impl(y, ldy, aux, ldaux, gy, ldgy, N, Fo, G, gpre, ldg, colsum, x, ldx, Fi, dw11, dw12, gate_dw_done, workspace,
                          workspace_bytes, stream);
}
}  // namespace gnnml3

extern "C" int gnnml3_ml3_act_bwd_y(const float* y, int64_t ldy, const float* aux, int64_t ldaux, const float* gy, int64_t ldgy,
                                    int64_t N, int Fo, int G, float* gpre, int64_t ldg, float* colsum, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
    return gnnml3::act_bwd_y_impl(y, ldy, aux, ldaux, gy, ldgy, N, Fo, G, gpre, ldg, colsum, nullptr, 0, 0, nullptr, nullptr, nullptr,
                                  workspace, workspace_bytes, stream_);
}

static int gnnml3::act_bwd_y_impl(const float* y, int64_t ldy, const float* aux, int64_t ldaux, const float* gy, int64_t ldgy, int64_t N,
                                  int Fo, int G, float* gpre, int64_t ldg, float* colsum, const float* x, int64_t ldx, int Fi,
                                  float* dw11, float* dw12, int* gate_dw_done, void* workspace, size_t workspace_bytes, void* stream_) {
    if (gate_dw_done) *gate_dw_done = 0;
    GNNML3_REQUIRE(N >= 0 && Fo >= 1 && G >= 0, "ml3_act_bwd_y: bad shape");
    cudaStream_t st = (cudaStream_t)stream_;
    if (N == 0) {
        if (colsum) GNNML3_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * (Fo + 2 * G), st));
        return GNNML3_OK;
    }
    const int Fo4 = (Fo + 3) / 4 * 4;
    GNNML3_REQUIRE(y && gy && gpre && ldy >= Fo && ldgy >= Fo + G && ldg >= Fo4 + 2 * G && ldg <= ACTY_MAXLD,
                   "ml3_act_bwd_y: bad arguments");
    GNNML3_REQUIRE(G == 0 || (aux && ldaux >= 2 * G), "ml3_act_bwd_y: aux [N, 2G] required");
    float* part = nullptr;
    if (colsum) {
        GNNML3_REQUIRE(workspace, "ml3_act_bwd_y: workspace required for the column sums");
        if (workspace_bytes < gnnml3_ml3_act_bwd_workspace_bytes(N, Fo, G))
            return set_err(GNNML3_ERR_WORKSPACE, "ml3_act_bwd_y: workspace too small");
        part = (float*)workspace;
    }
    GNNML3_REQUIRE(ldg % 4 == 0 && (uintptr_t)gpre % 16 == 0, "ml3_act_bwd_y: gpre rows must be 16-byte aligned");
    const int nb = cdiv(N, ACTY_ROWS);
    const int vec_in = (ldy % 4 == 0 && ldgy % 4 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)gy % 16 == 0) ? 1 : 0;
    // streaming kernel for the aligned layer shapes (every GNNML3 configuration); the general kernel for the rest
    const bool gates_ok = G == 0 || (G == 2 && ldaux % 4 == 0 && ((uintptr_t)aux & 15) == 0) ||
                          (G % 4 == 0 && (G & (G - 1)) == 0 && G <= 32 && Fo % 4 == 0 && ldaux % 4 == 0 && ((uintptr_t)aux & 15) == 0);
    static const bool general_only = [] { const char* e = getenv("GNNML3_ACT_GENERAL"); return e && e[0] == '1'; }();   // measurement switch
    const bool fast = !general_only && vec_in && gates_ok && ldg == Fo4 + 2 * G && (Fo4 == 16 || Fo4 == 32 || Fo4 == 64) && ldy >= Fo4 && ldgy >= Fo4;
    const int W2 = Fo + 2 * G;
    static const bool no_gw = [] { const char* e = getenv("GNNML3_ACT_NO_GATE_DW"); return e && e[0] == '1'; }();      // measurement switch
    const bool gw = fast && !no_gw && G == 2 && x && dw11 && dw12 && part && Fi >= 1 && Fi <= 32 && ldx % 4 == 0 && ldx >= (Fi + 3) / 4 * 4 &&
                    ((uintptr_t)x & 15) == 0;
    if (gw) {
        const int ps = W2 + ACTY_GW_COLS;
        if (Fo4 == 16) k_ml3_act_bwd_y_v<4, true><<<nb, 256, 0, st>>>(y, ldy, aux, ldaux, gy, ldgy, N, Fo, G, gpre, ldg, part, ps, x, ldx, Fi);
        else if (Fo4 == 32) k_ml3_act_bwd_y_v<8, true><<<nb, 256, 0, st>>>(y, ldy, aux, ldaux, gy, ldgy, N, Fo, G, gpre, ldg, part, ps, x, ldx, Fi);
        else k_ml3_act_bwd_y_v<16, true><<<nb, 256, 0, st>>>(y, ldy, aux, ldaux, gy, ldgy, N, Fo, G, gpre, ldg, part, ps, x, ldx, Fi);
        GNNML3_LAUNCH_CHECK();
        k_colsum_finish_gw<<<ps, 128, 0, st>>>(part, nb, ps, W2, Fi, colsum, dw11, dw12);
        GNNML3_LAUNCH_CHECK();
        if (gate_dw_done) *gate_dw_done = 1;
        return GNNML3_OK;
    }
    if (fast && Fo4 == 16) k_ml3_act_bwd_y_v<4, false><<<nb, 256, 0, st>>>(y, ldy, aux, ldaux, gy, ldgy, N, Fo, G, gpre, ldg, part, W2, nullptr, 0, 0);
    else if (fast && Fo4 == 32) k_ml3_act_bwd_y_v<8, false><<<nb, 256, 0, st>>>(y, ldy, aux, ldaux, gy, ldgy, N, Fo, G, gpre, ldg, part, W2, nullptr, 0, 0);
    else if (fast) k_ml3_act_bwd_y_v<16, false><<<nb, 256, 0, st>>>(y, ldy, aux, ldaux, gy, ldgy, N, Fo, G, gpre, ldg, part, W2, nullptr, 0, 0);
    else k_ml3_act_bwd_y<<<nb, 256, 0, st>>>(y, ldy, aux, ldaux, gy, ldgy, N, Fo, G, gpre, ldg, part, vec_in);
    GNNML3_LAUNCH_CHECK();
    if (colsum) {
        k_colsum_finish<<<cdiv(W2, 8), 256, 0, st>>>(part, nb, W2, colsum);
        GNNML3_LAUNCH_CHECK();
    }
    return GNNML3_OK;
}
