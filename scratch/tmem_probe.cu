// Round-2 probe (run on a B200 box): facts the TS-mode (A operand in tensor memory) fused layer kernel is built on.
//   A. tcgen05.st fragment layouts (16x64b.x16, 16x256b.x4, 16x128b.x8) read back through the known 32x32b.x32 load
//   B. tcgen05.mma kind::tf32 with A in TMEM (M = 128, K = 8 per instruction, B = K-major SWIZZLE_128B plane in smem):
//      numerics against a CPU product with truncated / rounded TF32 inputs
//   C. cycles per tcgen05.mma as a function of N, M and the operand source, with a tight unrolled issue loop (the round-1
//      probe divided by a runtime value inside the loop and measured its own loop: 142 cycles whatever the shape)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../gnn_matlang_b200/csrc -o tmem_probe tmem_probe.cu
#include "tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
using namespace gnnml3;

namespace gnnml3 {
char* err_buf() { static char b[256]; return b; }
int set_err(int code, const char*, ...) { return code; }
void count_launch(int) {}
}


__device__ __forceinline__ void st_16x64b_x16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.16x64b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void st_16x256b_x4(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void st_16x128b_x8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void st_32x32b_x32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ------------------------------------------------------------------------------------------------ A: store layouts
// out[shape][lane 0..127][col 0..31] = value found; the value written by thread T of warp w, register j, half h is
// 100000*h + 1000*T + j  (h = 1: the second store of the 16-lane shapes at lane offset 16)
__global__ void __launch_bounds__(128, 1) k_layout(float* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    for (int shape = 0; shape < 3; ++shape) {
        uint32_t r[16];
        for (int h = 0; h < 2; ++h) {
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint((float)(100000 * h + 1000 * lane + j));
            const uint32_t taddr = tm + ((uint32_t)(warp * 32 + 16 * h) << 16) + shape * 32;
            if (shape == 0) st_16x64b_x16(taddr, r);
            if (shape == 1) st_16x256b_x4(taddr, r);
            if (shape == 2) st_16x128b_x8(taddr, r);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    for (int shape = 0; shape < 3; ++shape) {
        float v[32];
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + shape * 32, v);
        for (int c = 0; c < 32; ++c) out[(shape * 128 + warp * 32 + lane) * 32 + c] = v[c];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 128);
}

// ------------------------------------------------------------------------------------------------ B: TS-mode numerics
// A [128 x 32] (row = TMEM lane) written with 32x32b.x32 (mode 0) or 16x64b.x16 with the decoded layout (mode 1);
// Bp [N=64 x 32] K-major SWIZZLE_128B plane; D[128 x 64] = A * Bp^T with 4 MMAs (N = 64), then 4 more MMAs with N = 32
// accumulating A2 * Bp[0:32]^T into columns 0..31.
__global__ void __launch_bounds__(128, 1) k_ts(const float* A, const float* A2, const float* Bp, float* D, int mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // B plane: row n (0..63), 16-byte chunk c (0..7) at n*128 + ((c ^ (n & 7)) << 4)
    for (int i = threadIdx.x; i < 64 * 32; i += 128) {
        const int n = i / 32, k = i % 32, c = k / 4;
        *reinterpret_cast<float*>(smem + n * 128 + ((c ^ (n & 7)) << 4) + (k % 4) * 4) = Bp[i];
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&slot, 256);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t colA = 128, colA2 = 160;           // A at columns 128..159, A2 at 160..191, D at 0..63
    if (mode == 0) {
        uint32_t r[32];
        const int row = warp * 32 + lane;
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(A[row * 32 + j]);
        st_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + colA, r);
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(A2[row * 32 + j]);
        st_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + colA2, r);
    } else {
        // 16x64b.x16: thread T -> row 8*(T&1) + (T>>2), columns ((T>>1)&1) + 2j
        for (int h = 0; h < 2; ++h) {
            uint32_t r[16];
            const int row = warp * 32 + 16 * h + 8 * (lane & 1) + (lane >> 2), p = (lane >> 1) & 1;
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(A[row * 32 + p + 2 * j]);
            st_16x64b_x16(tm + ((uint32_t)(warp * 32 + 16 * h) << 16) + colA, r);
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(A2[row * 32 + p + 2 * j]);
            st_16x64b_x16(tm + ((uint32_t)(warp * 32 + 16 * h) << 16) + colA2, r);
        }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem));
        const uint32_t id64 = make_idesc_tf32_mn(128, 64), id32 = make_idesc_tf32_mn(128, 32);
        for (int k = 0; k < 4; ++k) umma_tf32_ts(tm, tm + colA + 8 * k, db + (uint64_t)((k * 32) >> 4), id64, k ? 1u : 0u);
        for (int k = 0; k < 4; ++k) umma_tf32_ts(tm, tm + colA2 + 8 * k, db + (uint64_t)((k * 32) >> 4), id32, 1u);
        umma_commit(&bar);
        mbar_wait(&bar, 0);
    }
    __syncthreads();
    tc_fence_after();
    float v[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int c = 0; c < 32; ++c) D[(warp * 32 + lane) * 64 + c0 + c] = v[c];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}

// ------------------------------------------------------------------------------------------------ C: MMA pacing
// src 0: A and B from shared memory (SS); src 1: A from tensor memory (TS).  pair != 0: alternate N and N/2 (the 3xTF32
// pair of the fused kernel).  64 MMAs per loop iteration, fully unrolled, fixed descriptors.
template <int SRC>
__global__ void __launch_bounds__(128, 1) k_pace(int M, int N, int pair, int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t id1 = make_idesc_tf32_mn(M, N), id2 = make_idesc_tf32_mn(M, pair ? N / 2 : N);
        const uint64_t da = make_kmajor_sw128_desc(smem_u32(smem));
        const uint64_t db = make_kmajor_sw128_desc(smem_u32(smem + 16384));
        const uint32_t ta = tm + 256;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                if (SRC == 0) {
                    umma_tf32(tm, da + (uint64_t)((u & 3) * 2), db + (uint64_t)((u & 3) * 2), id1, 1u);
                    umma_tf32(tm, da + (uint64_t)((u & 3) * 2), db + (uint64_t)((u & 3) * 2), id2, 1u);
                } else {
                    umma_tf32_ts(tm, ta + 8 * (u & 3), db + (uint64_t)((u & 3) * 2), id1, 1u);
                    umma_tf32_ts(tm, ta + 32 + 8 * (u & 3), db + (uint64_t)((u & 3) * 2), id2, 1u);
                }
            }
        }
        umma_commit(&bar);
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

static float tf32_trunc(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    u &= 0xFFFFE000u;
    memcpy(&v, &u, 4);
    return v;
}
static float tf32_round(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    u += 0x1000u;
    u &= 0xFFFFE000u;
    memcpy(&v, &u, 4);
    return v;
}

int main() {
    // ---------------- A
    {
        float* d;
        cudaMalloc(&d, 3 * 128 * 32 * 4);
        cudaMemset(d, 0, 3 * 128 * 32 * 4);
        k_layout<<<1, 128>>>(d);
        cudaError_t e = cudaDeviceSynchronize();
        printf("A: layout kernel: %s\n", cudaGetErrorString(e));
        std::vector<float> h(3 * 128 * 32);
        cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
        const char* names[3] = {"16x64b.x16", "16x256b.x4", "16x128b.x8"};
        for (int s = 0; s < 3; ++s) {
            printf("A: shape %s, warp 0 quarter: value = 100000*half + 1000*thread + reg\n", names[s]);
            for (int lane = 0; lane < 32; ++lane) {
                printf("  lane %2d:", lane);
                for (int c = 0; c < 32; ++c) printf(" %6d", (int)h[(s * 128 + lane) * 32 + c]);
                printf("\n");
            }
            // check the decoded layouts
            int bad = 0;
            for (int w = 0; w < 4; ++w)
                for (int lane = 0; lane < 32; ++lane)
                    for (int c = 0; c < 32; ++c) {
                        const int hlf = lane / 16, l16 = lane % 16;
                        int T = -1, j = -1;
                        if (s == 0) { T = (l16 / 8) + 2 * (c & 1) + 4 * (l16 % 8); j = c / 2; }
                        if (s == 1) { T = 4 * (l16 % 8) + (c % 8) / 2; j = (c & 1) + 2 * (l16 / 8) + 4 * (c / 8); }
                        if (s == 2) { T = 4 * (l16 % 8) + (c % 4); j = (l16 / 8) + 2 * (c / 4); }
                        const int expect = 100000 * hlf + 1000 * T + j;
                        if ((int)h[(s * 128 + w * 32 + lane) * 32 + c] != expect) ++bad;
                    }
            printf("A: shape %s decoded-layout mismatches: %d of 4096\n", names[s], bad);
        }
        cudaFree(d);
    }
    // ---------------- B
    {
        std::vector<float> A(128 * 32), A2(128 * 32), Bp(64 * 32);
        srand(1);
        auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
        for (auto& v : A) v = rnd();
        for (auto& v : A2) v = rnd() * 1e-3f;
        for (auto& v : Bp) v = rnd();
        float *dA, *dA2, *dB, *dD;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dA2, A2.size() * 4); cudaMalloc(&dB, Bp.size() * 4); cudaMalloc(&dD, 128 * 64 * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dA2, A2.data(), A2.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, Bp.data(), Bp.size() * 4, cudaMemcpyHostToDevice);
        cudaFuncSetAttribute(k_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000);
        for (int mode = 0; mode < 2; ++mode) {
            cudaMemset(dD, 0, 128 * 64 * 4);
            k_ts<<<1, 128, 20000>>>(dA, dA2, dB, dD, mode);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> D(128 * 64);
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double errT = 0, errR = 0, errF = 0, mx = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    double sT = 0, sR = 0, sF = 0;
                    for (int k = 0; k < 32; ++k) {
                        sT += (double)tf32_trunc(A[m * 32 + k]) * tf32_trunc(Bp[n * 32 + k]);
                        sR += (double)tf32_round(A[m * 32 + k]) * tf32_round(Bp[n * 32 + k]);
                        sF += (double)A[m * 32 + k] * Bp[n * 32 + k];
                        if (n < 32) {
                            sT += (double)tf32_trunc(A2[m * 32 + k]) * tf32_trunc(Bp[n * 32 + k]);
                            sR += (double)tf32_round(A2[m * 32 + k]) * tf32_round(Bp[n * 32 + k]);
                            sF += (double)A2[m * 32 + k] * Bp[n * 32 + k];
                        }
                    }
                    const double d = D[m * 64 + n];
                    errT = fmax(errT, fabs(d - sT)); errR = fmax(errR, fabs(d - sR)); errF = fmax(errF, fabs(d - sF));
                    mx = fmax(mx, fabs(sF));
                }
            printf("B: TS-mode MMA, A written with %s: %s; max|D| %.3f, max err vs truncated-TF32 inputs %.3e, vs rounded %.3e, vs FP32 %.3e\n",
                   mode == 0 ? "32x32b.x32" : "16x64b.x16 (decoded layout)", cudaGetErrorString(e), mx, errT, errR, errF);
        }
    }
    // ---------------- C
    {
        long long* d;
        cudaMalloc(&d, 16);
        cudaFuncSetAttribute(k_pace<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
        cudaFuncSetAttribute(k_pace<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
        const int iters = 16;
        for (int src = 0; src < 2; ++src)
            for (int M : {64, 128})
                for (int N : {16, 32, 64, 96, 128, 256})
                    for (int pair = 0; pair < 2; ++pair) {
                        if (src == 1 && M == 64) continue;
                        if (M == 128 && (N % 16 || (pair && (N / 2) % 16))) continue;
                        if (pair && (N / 2) % 8) continue;
                        long long h[2];
                        for (int rep = 0; rep < 2; ++rep) {
                            if (src == 0) k_pace<0><<<1, 128, 100000>>>(M, N, pair, iters, d);
                            else k_pace<1><<<1, 128, 100000>>>(M, N, pair, iters, d);
                            cudaError_t e = cudaDeviceSynchronize();
                            if (e != cudaSuccess) { printf("C: error %s\n", cudaGetErrorString(e)); return 1; }
                        }
                        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                        const double n = iters * 64.0;
                        printf("C: %s M=%3d N=%3d%s : issue %.1f cyc/mma, complete %.1f cyc/mma\n", src ? "TS" : "SS", M, N,
                               pair ? " (+N/2 pair)" : "", h[0] / n, h[1] / n);
                    }
    }
    return 0;
}
