// SpectralDesign on the GPU: one thread block per graph.
//
// Reference: libs/utils.py:546-610 -- a python loop over graphs, each doing a dense numpy eigh (LAPACK dsyevd)
// plus nfreq dense U diag(f_k) U^T products, on one CPU core.  Here every graph of a batch is handled by one
// CTA entirely in shared memory:
//   1. adjacency / receptive-field mask as bit rows (A, then (A+I)^(2^(r-1)) > 0 by repeated boolean squaring)
//   2. the symmetric matrix to decompose, built exactly as the reference does (normalised Laplacian from FP32
//      factors 1/sqrt(d) with inf -> 0, lower triangle as read by eigh; or the adjacency when laplacien=False)
//   3. cyclic Jacobi eigensolver in FP64 with round-robin parallel ordering (n/2 disjoint rotations per round)
//   4. band filters exp(-dv (lambda - c_k)^2), c_k = linspace(lambda_min, vmax or lambda_max, nfreq)
//   5. supports evaluated ONLY at the mask positions: S_k[i,j] = sum_m U[i,m] f_k(lambda_m) U[j,m], identity and
//      optional adjacency channels, emitted in the reference's row-major np.where order as
//      edge_index2 [2,E2] (int64) / edge_attr2 [E2, nfreq+1(+1)] (FP32).
// Supports are invariant to eigenvector sign / rotation inside degenerate eigenspaces, so they (not U) are what
// the parity tests compare.  A first pass (gnnml3_spectral_count) sizes the output.
#include "common.cuh"

#include <cmath>

namespace gnnml3 {

constexpr int SD_THREADS = 256;
constexpr int SD_WORDS = 4;      // 128-bit rows -> n <= 128 for the mask; the eigensolver's smem bounds n further
constexpr int SD_MAX_SWEEPS = 48;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < SD_THREADS / 32; ++w) s += red[w];
    return s;
}

// adjacency bits + receptive-field mask bits for one graph; returns the buffer index holding the mask
__device__ int build_mask(const int64_t* __restrict__ ei, int64_t Etot, int e0, int e1, int n, int recfield,
                          uint32_t* Abits, uint32_t* M0, uint32_t* M1) {
    for (int i = threadIdx.x; i < n * SD_WORDS; i += blockDim.x) Abits[i] = 0u;
    __syncthreads();
    for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int s = (int)ei[e], t = (int)ei[Etot + e];
        if (s >= 0 && s < n && t >= 0 && t < n) atomicOr(&Abits[s * SD_WORDS + (t >> 5)], 1u << (t & 31));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * SD_WORDS; i += blockDim.x) {
        uint32_t v = Abits[i];
        const int r = i / SD_WORDS, w = i % SD_WORDS;
        if (recfield > 0 && (r >> 5) == w) v |= 1u << (r & 31);
        M0[i] = v;
    }
    __syncthreads();
    uint32_t* cur = M0;
    uint32_t* nxt = M1;
    for (int it = 1; it < recfield; ++it) {          // M <- (M M) > 0
        for (int i = threadIdx.x; i < n * SD_WORDS; i += blockDim.x) {
            const int r = i / SD_WORDS, w = i % SD_WORDS;
            uint32_t acc = 0u;
            for (int ww = 0; ww < SD_WORDS; ++ww) {
                uint32_t bits = cur[r * SD_WORDS + ww];
                while (bits) {
                    const int j = (ww << 5) + __ffs(bits) - 1;
                    bits &= bits - 1;
                    acc |= cur[j * SD_WORDS + w];
                }
            }
            nxt[i] = acc;
        }
        __syncthreads();
        uint32_t* t = cur; cur = nxt; nxt = t;
    }
    return cur == M0 ? 0 : 1;
}

__global__ void __launch_bounds__(SD_THREADS)
k_sd_count(const int64_t* __restrict__ ei, int64_t Etot, const int* __restrict__ edge_ptr, const int* __restrict__ node_ptr,
           int recfield, int* __restrict__ counts) {
    extern __shared__ __align__(16) uint32_t sbits[];
    const int b = blockIdx.x;
    const int n = node_ptr[b + 1] - node_ptr[b];
    uint32_t* Abits = sbits;
    uint32_t* M0 = Abits + n * SD_WORDS;
    uint32_t* M1 = M0 + n * SD_WORDS;
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    const int which = build_mask(ei, Etot, edge_ptr[b], edge_ptr[b + 1], n, recfield, Abits, M0, M1);
    const uint32_t* M = which ? M1 : M0;
    int c = 0;
    for (int i = threadIdx.x; i < n * SD_WORDS; i += blockDim.x) c += __popc(M[i]);
    atomicAdd(&total, c);
    __syncthreads();
    if (threadIdx.x == 0) counts[b] = total;
}

// cyclic Jacobi on the symmetric n x n matrix in As (leading dimension ld, odd); eigenvectors accumulate in Us
__device__ void jacobi_eigh(double* As, double* Us, int n, int ld, double* cs, double* red) {
    const int tid = threadIdx.x;
    for (int i = tid; i < n * n; i += blockDim.x) {
        const int r = i / n, c = i % n;
        Us[r * ld + c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (n < 2) return;
    const int m = (n + 1) & ~1;            // players of the round-robin tournament (dummy index n if n is odd)
    const int half = m / 2;
    for (int sweep = 0; sweep < SD_MAX_SWEEPS; ++sweep) {
        double off = 0.0, all = 0.0;
        for (int i = tid; i < n * n; i += blockDim.x) {
            const int r = i / n, c = i % n;
            const double v = As[r * ld + c];
            all += v * v;
            if (r != c) off += v * v;
        }
        off = block_sum(off, red);
        all = block_sum(all, red);
        if (off <= 1e-28 * all || all == 0.0) break;
        for (int round = 0; round < m - 1; ++round) {
            // 1. rotation angles of the round's disjoint pairs
            if (tid < half) {
                int p, q;
                if (tid == 0) { p = m - 1; q = round % (m - 1); }
                else { p = (round + tid) % (m - 1); q = (round - tid + (m - 1)) % (m - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                double c = 1.0, s = 0.0;
                if (q < n) {
                    const double apq = As[p * ld + q];
                    if (apq != 0.0) {
                        // t = sign(theta) / (|theta| + sqrt(theta^2 + 1)) with theta = d / b, d = a_qq - a_pp, b = 2 a_pq, written
                        // without forming theta:  t = sign(d b) |b| / (|d| + sqrt(d^2 + b^2)) -- one sqrt, one division and one
                        // rsqrt on this serial chain of the round instead of three divisions and two square roots (the chain,
                        // not the rotations, was the longest part of a round), and no overflow for tiny a_pq
                        const double d = As[q * ld + q] - As[p * ld + p], b = 2.0 * apq;
                        const double t = (((d >= 0.0) == (b >= 0.0)) ? 1.0 : -1.0) * fabs(b) / (fabs(d) + sqrt(d * d + b * b));
                        c = rsqrt(t * t + 1.0);
                        s = t * c;
                    }
                }
                cs[4 * tid + 0] = c;
                cs[4 * tid + 1] = s;
                cs[4 * tid + 2] = (double)p;
                cs[4 * tid + 3] = (double)q;
            }
            __syncthreads();
            // 2. column rotations of A and U:  X[:, p], X[:, q] <- X[:, p] c - X[:, q] s,  X[:, p] s + X[:, q] c
            for (int i = tid; i < half * n; i += blockDim.x) {
                const int k = i / n, r = i % n;
                const int p = (int)cs[4 * k + 2], q = (int)cs[4 * k + 3];
                if (q >= n) continue;
                const double c = cs[4 * k], s = cs[4 * k + 1];
                if (s == 0.0) continue;
                const double ap = As[r * ld + p], aq = As[r * ld + q];
                As[r * ld + p] = c * ap - s * aq;
                As[r * ld + q] = s * ap + c * aq;
                const double up = Us[r * ld + p], uq = Us[r * ld + q];
                Us[r * ld + p] = c * up - s * uq;
                Us[r * ld + q] = s * up + c * uq;
            }
            __syncthreads();
            // 3. row rotations of A
            for (int i = tid; i < half * n; i += blockDim.x) {
                const int k = i / n, col = i % n;
                const int p = (int)cs[4 * k + 2], q = (int)cs[4 * k + 3];
                if (q >= n) continue;
                const double c = cs[4 * k], s = cs[4 * k + 1];
                if (s == 0.0) continue;
                const double ap = As[p * ld + col], aq = As[q * ld + col];
                As[p * ld + col] = c * ap - s * aq;
                As[q * ld + col] = s * ap + c * aq;
            }
            __syncthreads();
        }
    }
}

struct SDParams {
    int recfield, nfreq, laplacien, addadj, has_vmax, global_ids;
    double dv, vmax;
};

__global__ void __launch_bounds__(SD_THREADS)
k_sd_design(const int64_t* __restrict__ ei, int64_t Etot, const int* __restrict__ edge_ptr, const int* __restrict__ node_ptr,
            SDParams P, int nmax, const int64_t* __restrict__ out_ptr, int64_t* __restrict__ ei2, int64_t E2,
            float* __restrict__ ea2, float* __restrict__ lmax_out, float* __restrict__ deg_out) {
    extern __shared__ __align__(16) unsigned char sraw[];
    const int b = blockIdx.x;
    const int nbeg = node_ptr[b];
    const int n = node_ptr[b + 1] - nbeg;
    const int ld = nmax | 1;
    const int K = P.nfreq + 1 + (P.addadj ? 1 : 0);
    // carve shared memory (sized for nmax by the host)
    double* As = reinterpret_cast<double*>(sraw);
    double* Us = As + (size_t)nmax * ld;
    double* lam = Us + (size_t)nmax * ld;
    double* ftab = lam + nmax;                        // [nfreq][nmax]
    double* cs = ftab + (size_t)P.nfreq * nmax;       // [nmax/2 + 1][4]
    double* red = cs + 4 * (nmax / 2 + 1);            // [8]
    float* dis = reinterpret_cast<float*>(red + 8);   // [nmax]
    int* rowoff = reinterpret_cast<int*>(dis + nmax); // [nmax + 1]
    uint32_t* Abits = reinterpret_cast<uint32_t*>(rowoff + nmax + 1);
    uint32_t* M0 = Abits + nmax * SD_WORDS;
    uint32_t* M1 = M0 + nmax * SD_WORDS;
    __shared__ double s_vmin, s_vmax;
    const int tid = threadIdx.x;
    if (n <= 0) return;

    const int which = build_mask(ei, Etot, edge_ptr[b], edge_ptr[b + 1], n, P.recfield, Abits, M0, M1);
    const uint32_t* M = which ? M1 : M0;

    // in-degree d = A.sum(axis=0) (libs/utils.py:562,576) and dis = 1/sqrt(d) in FP32 with inf -> 0 (:578-580)
    for (int j = tid; j < n; j += blockDim.x) {
        int d = 0;
        for (int i = 0; i < n; ++i) d += (Abits[i * SD_WORDS + (j >> 5)] >> (j & 31)) & 1u;
        dis[j] = d > 0 ? 1.0f / sqrtf((float)d) : 0.f;
        if (deg_out) deg_out[nbeg + j] = (float)d;
    }
    if (tid == 0) {                                    // row offsets of the emitted entries
        int acc = 0;
        for (int i = 0; i < n; ++i) {
            rowoff[i] = acc;
            for (int w = 0; w < SD_WORDS; ++w) acc += __popc(M[i * SD_WORDS + w]);
        }
        rowoff[n] = acc;
    }
    __syncthreads();

    // normalised Laplacian, lower triangle mirrored (np.linalg.eigh reads 'L'): nL[i][j] = delta - A[j][i] dis_i dis_j
    for (int i = tid; i < n * n; i += blockDim.x) {
        const int r = i / n, c = i % n;
        const int hi = r > c ? r : c, lo = r > c ? c : r;      // entry (hi, lo) of the lower triangle
        double v = (r == c) ? 1.0 : 0.0;
        const uint32_t a = (Abits[lo * SD_WORDS + (hi >> 5)] >> (hi & 31)) & 1u;   // A[lo][hi] = A[j][i] with i=hi, j=lo
        if (a) v -= (double)__fmul_rn(__fmul_rn(1.0f, dis[hi]), dis[lo]);
        As[r * ld + c] = v;
    }
    __syncthreads();
    jacobi_eigh(As, Us, n, ld, cs, red);
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) lam[i] = fmax(As[i * ld + i], 0.0);      // V[V<0] = 0 (:584)
    __syncthreads();
    if (tid == 0) {
        double mx = lam[0];
        for (int i = 1; i < n; ++i) mx = fmax(mx, lam[i]);
        if (lmax_out) lmax_out[b] = (float)mx;
    }
    if (!P.laplacien) {                                 // :588-589  eigh(A), eigenvalues NOT clamped
        __syncthreads();
        for (int i = tid; i < n * n; i += blockDim.x) {
            const int r = i / n, c = i % n;
            const int hi = r > c ? r : c, lo = r > c ? c : r;
            As[r * ld + c] = (double)((Abits[hi * SD_WORDS + (lo >> 5)] >> (lo & 31)) & 1u);    // A[hi][lo]
        }
        __syncthreads();
        jacobi_eigh(As, Us, n, ld, cs, red);
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) lam[i] = As[i * ld + i];
    }
    __syncthreads();
    if (tid == 0) {
        double mn = lam[0], mx = lam[0];
        for (int i = 1; i < n; ++i) { mn = fmin(mn, lam[i]); mx = fmax(mx, lam[i]); }
        s_vmin = mn;
        s_vmax = P.has_vmax ? P.vmax : mx;
    }
    __syncthreads();
    // band filters (:592-600): centers = linspace(vmin, vmax, nfreq)
    for (int i = tid; i < P.nfreq * n; i += blockDim.x) {
        const int k = i / n, mm = i % n;
        const double step = P.nfreq > 1 ? (s_vmax - s_vmin) / (double)(P.nfreq - 1) : 0.0;
        const double c = (k == P.nfreq - 1 && P.nfreq > 1) ? s_vmax : s_vmin + step * k;
        const double d = lam[mm] - c;
        ftab[k * nmax + mm] = exp(-(P.dv * d * d));
    }
    __syncthreads();

    // supports at the mask positions, row-major order
    const int64_t base = out_ptr[b];
    const int64_t goff = P.global_ids ? (int64_t)nbeg : 0;
    for (int i = tid; i < n * n; i += blockDim.x) {
        const int r = i / n, c = i % n;
        if (!((M[r * SD_WORDS + (c >> 5)] >> (c & 31)) & 1u)) continue;
        int pos = rowoff[r];
        for (int w = 0; w < (c >> 5); ++w) pos += __popc(M[r * SD_WORDS + w]);
        pos += __popc(M[r * SD_WORDS + (c >> 5)] & ((1u << (c & 31)) - 1u));
        const int64_t o = base + pos;
        ei2[o] = goff + r;
        ei2[E2 + o] = goff + c;
        float* dst = ea2 + o * K;
        const double* ur = Us + (size_t)r * ld;
        const double* uc = Us + (size_t)c * ld;
        for (int k0 = 0; k0 < P.nfreq; k0 += 4) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            const int kn = min(4, P.nfreq - k0);
            for (int mm = 0; mm < n; ++mm) {
                const double pr = ur[mm] * uc[mm];
                a0 += pr * ftab[(k0 + 0) * nmax + mm];
                if (kn > 1) a1 += pr * ftab[(k0 + 1) * nmax + mm];
                if (kn > 2) a2 += pr * ftab[(k0 + 2) * nmax + mm];
                if (kn > 3) a3 += pr * ftab[(k0 + 3) * nmax + mm];
            }
            dst[k0] = (float)a0;
            if (kn > 1) dst[k0 + 1] = (float)a1;
            if (kn > 2) dst[k0 + 2] = (float)a2;
            if (kn > 3) dst[k0 + 3] = (float)a3;
        }
        dst[P.nfreq] = (r == c) ? 1.f : 0.f;                                           // :602 identity
        if (P.addadj) dst[P.nfreq + 1] = (float)((Abits[r * SD_WORDS + (c >> 5)] >> (c & 31)) & 1u);   // :604-605
    }
}

static size_t sd_smem_bytes(int nmax, int nfreq) {
    const size_t ld = (size_t)(nmax | 1);
    size_t dbl = 2 * (size_t)nmax * ld + nmax + (size_t)nfreq * nmax + 4 * (nmax / 2 + 1) + 8;
    size_t bytes = dbl * sizeof(double) + (size_t)nmax * sizeof(float) + (size_t)(nmax + 1) * sizeof(int) +
                   3 * (size_t)nmax * SD_WORDS * sizeof(uint32_t);
    return align_up(bytes, 16);
}

}  // namespace gnnml3

using namespace gnnml3;

extern "C" int gnnml3_spectral_max_nodes(int nfreq) {
    int n = 128;
    while (n > 1 && sd_smem_bytes(n, nfreq) > 227 * 1024) --n;
    return n;
}

extern "C" int gnnml3_spectral_count(const int64_t* edge_index, int64_t Etot, const int32_t* edge_ptr, const int32_t* node_ptr,
                                     int B, int recfield, int nmax, int32_t* counts, void* stream_) {
    GNNML3_REQUIRE(B >= 0 && Etot >= 0 && recfield >= 0, "spectral_count: bad arguments");
    if (B == 0) return GNNML3_OK;
    GNNML3_REQUIRE(edge_ptr && node_ptr && counts && (Etot == 0 || edge_index), "spectral_count: NULL pointer");
    GNNML3_REQUIRE(nmax >= 1 && nmax <= 128, "spectral_count: graphs of up to 128 nodes are supported (nmax=%d)", nmax);
    const size_t smem = 3 * (size_t)nmax * SD_WORDS * sizeof(uint32_t);
    k_sd_count<<<B, SD_THREADS, smem, (cudaStream_t)stream_>>>(edge_index, Etot, edge_ptr, node_ptr, recfield, counts);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

extern "C" int gnnml3_spectral_design(const int64_t* edge_index, int64_t Etot, const int32_t* edge_ptr, const int32_t* node_ptr,
                                      int B, int recfield, double dv, int nfreq, int laplacien, int addadj, int has_vmax,
                                      double vmax, int nmax, const int64_t* out_ptr, int global_ids, int64_t* edge_index2,
                                      int64_t E2, float* edge_attr2, float* lmax, float* degree, void* stream_) {
    GNNML3_REQUIRE(B >= 0 && Etot >= 0 && recfield >= 0 && nfreq >= 1, "spectral_design: bad arguments");
    if (B == 0) return GNNML3_OK;
    GNNML3_REQUIRE(edge_ptr && node_ptr && out_ptr && (Etot == 0 || edge_index), "spectral_design: NULL pointer");
    GNNML3_REQUIRE(E2 == 0 || (edge_index2 && edge_attr2), "spectral_design: NULL output");
    GNNML3_REQUIRE(nmax >= 1 && nmax <= gnnml3_spectral_max_nodes(nfreq),
                   "spectral_design: largest graph has %d nodes; the one-block-per-graph eigensolver holds up to %d "
                   "(nfreq=%d) in shared memory", nmax, gnnml3_spectral_max_nodes(nfreq), nfreq);
    SDParams P;
    P.recfield = recfield; P.nfreq = nfreq; P.laplacien = laplacien; P.addadj = addadj; P.has_vmax = has_vmax;
    P.global_ids = global_ids; P.dv = dv; P.vmax = vmax;
    const size_t smem = sd_smem_bytes(nmax, nfreq);
    {   // the attribute grows with the largest graph seen on this device; guarded like every other per-device configuration
        static size_t configured[64] = {};
        int dev_ = 0;
        cudaGetDevice(&dev_);
        dev_ &= 63;
        std::lock_guard<std::mutex> lk(config_mutex());
        if (smem > configured[dev_]) {
            GNNML3_CUDA(cudaFuncSetAttribute(k_sd_design, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[dev_] = smem;
        }
    }
    k_sd_design<<<B, SD_THREADS, smem, (cudaStream_t)stream_>>>(edge_index, Etot, edge_ptr, node_ptr, P, nmax, out_ptr,
                                                                 edge_index2, E2, edge_attr2, lmax, degree);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}
