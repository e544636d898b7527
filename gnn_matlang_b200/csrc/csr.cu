// dst-sorted CSR and src-sorted (transposed) CSR of the batched disjoint graph -- bit-exact integer work.
//
// Replaces the per-support index_select / scatter_add indexing of PyG's MessagePassing.propagate that
// SpectConv.forward calls K times (reference libs/spect_conv.py:77): the batch's edge_index2 is sorted
// once by target (stable: entries of a row keep the original edge order, which is the order in which the
// reference's CPU scatter_add accumulates) and every layer of the model reuses it.
//
// Pipeline (all int32, E,N < 2^31): histogram -> exclusive scan -> atomic fill (arbitrary order inside a
// row) -> per-row rank sort by original edge id (makes the result deterministic and stable).
#include "common.cuh"

namespace gnnml3 {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void k_csr_hist(const int64_t* __restrict__ ei, int64_t E, int N, int* __restrict__ cnt_dst,
                           int* __restrict__ cnt_src, int* __restrict__ err) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = ei[e], t = ei[E + e];
        if (s < 0 || s >= N || t < 0 || t >= N) {
            if (err) *err = 1;
            s = min(max(s, (int64_t)0), (int64_t)N - 1);
            t = min(max(t, (int64_t)0), (int64_t)N - 1);
        }
        atomicAdd(cnt_dst + t, 1);
        atomicAdd(cnt_src + s, 1);
    }
}

// block-wide exclusive scan of kScanTile ints held kScanItems per thread (blocked arrangement)
__device__ __forceinline__ int block_exclusive_scan(int (&v)[kScanItems], int* smem_warp /*[8]*/) {
    int tsum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int x = v[i];
        v[i] = tsum;
        tsum += x;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) smem_warp[warp] = incl;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        int s = smem_warp[w];
        if (w < warp) woff += s;
        total += s;
    }
    const int toff = woff + incl - tsum;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) v[i] += toff;
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_sums(const int* __restrict__ in, int n, int* __restrict__ bsum) {
    __shared__ int sw[kScanThreads / 32];
    int v[kScanItems];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) v[i] = (base + i < n) ? in[base + i] : 0;
    int total = block_exclusive_scan(v, sw);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// single block: exclusive scan of the per-tile sums (nb may exceed one tile -> sequential chunks)
__global__ void __launch_bounds__(kScanThreads) k_scan_bsums(int* __restrict__ bsum, int nb) {
    __shared__ int sw[kScanThreads / 32];
    int carry = 0;
    for (int c0 = 0; c0 < nb; c0 += kScanTile) {
        int v[kScanItems];
        const int base = c0 + threadIdx.x * kScanItems;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) v[i] = (base + i < nb) ? bsum[base + i] : 0;
        int total = block_exclusive_scan(v, sw);
#pragma unroll
        for (int i = 0; i < kScanItems; ++i)
            if (base + i < nb) bsum[base + i] = v[i] + carry;
        carry += total;
    }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_final(const int* __restrict__ in, int n, const int* __restrict__ bsum,
                                                             int* __restrict__ out, int* __restrict__ out2) {
    __shared__ int sw[kScanThreads / 32];
    int v[kScanItems];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) v[i] = (base + i < n) ? in[base + i] : 0;
    block_exclusive_scan(v, sw);
    const int off = bsum[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) {
            out[base + i] = v[i] + off;
            out2[base + i] = v[i] + off;
        }
}

__global__ void k_csr_fill(const int64_t* __restrict__ ei, int64_t E, int N, int* __restrict__ cur_dst,
                           int* __restrict__ cur_src, int* __restrict__ tmp_dst, int* __restrict__ tmp_src) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = min(max(ei[e], (int64_t)0), (int64_t)N - 1);
        int64_t t = min(max(ei[E + e], (int64_t)0), (int64_t)N - 1);
        tmp_dst[atomicAdd(cur_dst + t, 1)] = (int)e;
        tmp_src[atomicAdd(cur_src + s, 1)] = (int)e;
    }
}

// rank sort inside each row: position p holds edge e of row r; its final slot is rowptr[r] + #{q in row: tmp[q] < e}.
// key_row selects which endpoint defines the row (1 = dst for the forward CSR, 0 = src for the transposed CSR);
// the other endpoint is written to `col`; `val` (optional) maps the edge id before it is written to `perm`.
__global__ void k_csr_rank(const int64_t* __restrict__ ei, int64_t E, int N, int key_row, const int* __restrict__ rowptr,
                           const int* __restrict__ tmp, const int* __restrict__ val, int* __restrict__ col,
                           int* __restrict__ perm) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < E; p += (int64_t)gridDim.x * blockDim.x) {
        const int e = tmp[p];
        int64_t kk = ei[(int64_t)key_row * E + e], oo = ei[(int64_t)(1 - key_row) * E + e];
        const int r = (int)min(max(kk, (int64_t)0), (int64_t)N - 1);
        const int o = (int)min(max(oo, (int64_t)0), (int64_t)N - 1);
        const int rs = rowptr[r], re = rowptr[r + 1];
        int rank = 0;
        for (int q = rs; q < re; ++q) rank += (tmp[q] < e) ? 1 : 0;
        col[rs + rank] = o;
        perm[rs + rank] = val ? val[e] : e;
    }
}

__global__ void k_invert_perm(const int* __restrict__ perm, int64_t E, int* __restrict__ inv) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < E; p += (int64_t)gridDim.x * blockDim.x)
        inv[perm[p]] = (int)p;
}

template <bool GATHER>
__global__ void k_permute_rows(const float* __restrict__ in, const int* __restrict__ perm, int64_t rows, int width,
                               float* __restrict__ out) {
    const int64_t total = rows * width;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / width;
        const int c = (int)(i - r * width);
        const int64_t pr = perm[r];
        if (GATHER)
            out[i] = __ldg(in + pr * width + c);
        else
            out[pr * width + c] = __ldg(in + i);
    }
}

// same permutation with 128-bit accesses (width % 4 == 0, 16-byte aligned buffers): one thread per 4 columns of a row
template <bool GATHER>
__global__ void k_permute_rows_v4(const float4* __restrict__ in, const int* __restrict__ perm, int64_t rows, int w4,
                                  float4* __restrict__ out) {
    const int64_t total = rows * w4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / w4;
        const int c = (int)(i - r * w4);
        const int64_t pr = __ldg(perm + r);
        if (GATHER)
            out[i] = __ldg(in + pr * w4 + c);
        else
            out[pr * w4 + c] = __ldg(in + i);
    }
}

static int grid_for(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace gnnml3

using namespace gnnml3;

extern "C" size_t gnnml3_csr_workspace_bytes(int64_t E, int64_t N) {
    const size_t nb = (size_t)cdiv(N + 1, kScanTile);
    size_t ints = 4 * (size_t)(N + 1) + 2 * (nb + 1) + 3 * (size_t)(E > 0 ? E : 1);
    return align_up(ints * sizeof(int), 256) + 256;
}

extern "C" int gnnml3_csr_build(const int64_t* edge_index, int64_t E, int64_t N, int32_t* rowptr, int32_t* col,
                                int32_t* perm, int32_t* rowptrT, int32_t* colT, int32_t* permT, int32_t* err_flag,
                                void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    GNNML3_REQUIRE(N > 0 && E >= 0, "csr_build: need N > 0 and E >= 0 (got N=%lld E=%lld)", (long long)N, (long long)E);
    GNNML3_REQUIRE(N < (1ll << 31) - kScanTile && E < (1ll << 31) - 1, "csr_build: N, E must fit int32");
    GNNML3_REQUIRE(rowptr && rowptrT, "csr_build: NULL rowptr");
    GNNML3_REQUIRE(E == 0 || (edge_index && col && perm && colT && permT), "csr_build: NULL pointer");
    if (workspace_bytes < gnnml3_csr_workspace_bytes(E, N))
        return set_err(GNNML3_ERR_WORKSPACE, "csr_build: workspace %zu < %zu bytes", workspace_bytes,
                       gnnml3_csr_workspace_bytes(E, N));
    const int n1 = (int)N + 1;
    const int nb = cdiv(n1, kScanTile);
    int* w = (int*)workspace;
    int* cnt_dst = w;            w += n1;
    int* cnt_src = w;            w += n1;
    int* cur_dst = w;            w += n1;
    int* cur_src = w;            w += n1;
    int* bs_dst = w;             w += nb + 1;
    int* bs_src = w;             w += nb + 1;
    int* tmp_dst = w;            w += (E > 0 ? E : 1);
    int* tmp_src = w;            w += (E > 0 ? E : 1);
    int* inv = w;
    GNNML3_CUDA(cudaMemsetAsync(cnt_dst, 0, sizeof(int) * 2 * (size_t)n1, st));
    if (err_flag) GNNML3_CUDA(cudaMemsetAsync(err_flag, 0, sizeof(int), st));
    const int ge = grid_for(E, 256);
    if (E > 0) {
        k_csr_hist<<<ge, 256, 0, st>>>(edge_index, E, (int)N, cnt_dst, cnt_src, err_flag);
        GNNML3_LAUNCH_CHECK();
    }
    k_scan_sums<<<nb, kScanThreads, 0, st>>>(cnt_dst, n1, bs_dst);
    GNNML3_LAUNCH_CHECK();
    k_scan_sums<<<nb, kScanThreads, 0, st>>>(cnt_src, n1, bs_src);
    GNNML3_LAUNCH_CHECK();
    k_scan_bsums<<<1, kScanThreads, 0, st>>>(bs_dst, nb);
    GNNML3_LAUNCH_CHECK();
    k_scan_bsums<<<1, kScanThreads, 0, st>>>(bs_src, nb);
    GNNML3_LAUNCH_CHECK();
    k_scan_final<<<nb, kScanThreads, 0, st>>>(cnt_dst, n1, bs_dst, rowptr, cur_dst);
    GNNML3_LAUNCH_CHECK();
    k_scan_final<<<nb, kScanThreads, 0, st>>>(cnt_src, n1, bs_src, rowptrT, cur_src);
    GNNML3_LAUNCH_CHECK();
    if (E > 0) {
        k_csr_fill<<<ge, 256, 0, st>>>(edge_index, E, (int)N, cur_dst, cur_src, tmp_dst, tmp_src);
        GNNML3_LAUNCH_CHECK();
        k_csr_rank<<<ge, 256, 0, st>>>(edge_index, E, (int)N, 1, rowptr, tmp_dst, nullptr, col, perm);
        GNNML3_LAUNCH_CHECK();
        k_invert_perm<<<ge, 256, 0, st>>>(perm, E, inv);
        GNNML3_LAUNCH_CHECK();
        k_csr_rank<<<ge, 256, 0, st>>>(edge_index, E, (int)N, 0, rowptrT, tmp_src, inv, colT, permT);
        GNNML3_LAUNCH_CHECK();
    }
    return GNNML3_OK;
}

extern "C" int gnnml3_gather_rows(const float* in, const int32_t* perm, int64_t rows, int width, float* out, void* stream_) {
    GNNML3_REQUIRE(rows >= 0 && width > 0, "gather_rows: bad shape");
    if (rows == 0) return GNNML3_OK;
    GNNML3_REQUIRE(in && perm && out, "gather_rows: NULL pointer");
    if (width % 4 == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0)
        k_permute_rows_v4<true><<<grid_for(rows * (width / 4), 256), 256, 0, (cudaStream_t)stream_>>>(
            reinterpret_cast<const float4*>(in), perm, rows, width / 4, reinterpret_cast<float4*>(out));
    else
        k_permute_rows<true><<<grid_for(rows * width, 256), 256, 0, (cudaStream_t)stream_>>>(in, perm, rows, width, out);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

extern "C" int gnnml3_scatter_rows(const float* in, const int32_t* perm, int64_t rows, int width, float* out, void* stream_) {
    GNNML3_REQUIRE(rows >= 0 && width > 0, "scatter_rows: bad shape");
    if (rows == 0) return GNNML3_OK;
    GNNML3_REQUIRE(in && perm && out, "scatter_rows: NULL pointer");
    if (width % 4 == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0)
        k_permute_rows_v4<false><<<grid_for(rows * (width / 4), 256), 256, 0, (cudaStream_t)stream_>>>(
            reinterpret_cast<const float4*>(in), perm, rows, width / 4, reinterpret_cast<float4*>(out));
    else
        k_permute_rows<false><<<grid_for(rows * width, 256), 256, 0, (cudaStream_t)stream_>>>(in, perm, rows, width, out);
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}
