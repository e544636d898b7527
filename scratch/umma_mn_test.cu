// Functional check of MN-major tcgen05.mma kind::tf32 operands in the SWIZZLE_128B layout (M = 64, N = 64, K = 8 rows).
#include "tc_common.cuh"
#include <cstdio>
#include <vector>
using namespace gnnml3;

constexpr int ROWS = 80;   // plane rows (as in the fused kernel); chunk stride = ROWS * 128 bytes

__global__ void __launch_bounds__(128, 1) k_test(const float* A, const float* B, float* D, int lane_off, int r8sel, int mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint8_t* pa = smem;                       // 2 chunks x [ROWS x 128 B]
    uint8_t* pb = smem + 2 * ROWS * 128;
    // A[row][m], m < 64: chunk m/32, swizzled 16-byte chunk ((m%32)/4) ^ (row%8)
    for (int i = threadIdx.x; i < 2 * 2 * ROWS * 128 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < ROWS * 64; i += blockDim.x) {
        const int row = i / 64, m = i % 64;
        const bool a_k = (mode == 0 || mode == 4 || mode == 5), b_k = (mode == 0 || mode == 3 || mode == 5);
        if (false) {
            // K-major reference: plane row = m (64 rows), k = row index within the selected group of 8 (32 bytes)
            if (row / 8 == r8sel) {
                const int k = row % 8;
                const uint32_t off = (m / 8) * 1024 + (m % 8) * 128 + (((k / 4) ^ (m % 8)) << 4) + (k % 4) * 4;
                *reinterpret_cast<float*>(pa + off) = A[i];
                *reinterpret_cast<float*>(pb + off) = B[i];
            }
        }
        {
            const int k = row % 8;
            const uint32_t offk = (m / 8) * 1024 + (m % 8) * 128 + (((k / 4) ^ (m % 8)) << 4) + (k % 4) * 4;
            const uint32_t offm = (m / 32) * ROWS * 128 + (row / 8) * 1024 + (row % 8) * 128 + ((((m % 32) / 4) ^ (row % 8)) << 4) + (m % 4) * 4;
            if (a_k) { if (row / 8 == r8sel) *reinterpret_cast<float*>(pa + offk) = A[i]; } else *reinterpret_cast<float*>(pa + offm) = A[i];
            if (b_k) { if (row / 8 == r8sel) *reinterpret_cast<float*>(pb + offk) = B[i]; } else *reinterpret_cast<float*>(pb + offm) = B[i];
        }
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        uint32_t idesc = make_idesc_tf32_mn(64, 64, true, true);
        uint64_t da = make_mnmajor_sw128_desc(smem_u32(pa), ROWS * 128);
        uint64_t db = make_mnmajor_sw128_desc(smem_u32(pb), ROWS * 128);
        uint64_t adv = (uint64_t)((r8sel * 1024) >> 4);
        if (mode == 0) {
            idesc = make_idesc_tf32_mn(64, 64, false, false);
            da = make_kmajor_sw128_desc(smem_u32(pa));
            db = make_kmajor_sw128_desc(smem_u32(pb));
            adv = 0;
        } else if (mode == 2) {          // LBO / SBO swapped
            auto mk = [](uint32_t addr, uint32_t lbo, uint32_t sbo) {
                uint64_t d = 0;
                d |= (uint64_t)((addr & 0x3FFFF) >> 4);
                d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
                d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
                d |= (uint64_t)1 << 46;
                d |= (uint64_t)2 << 61;
                return d;
            };
            da = mk(smem_u32(pa), 1024, ROWS * 128);
            db = mk(smem_u32(pb), 1024, ROWS * 128);
        }
        uint64_t adva = adv, advb = adv;
        if (mode >= 3) {
            const bool a_k = (mode == 4 || mode == 5), b_k = (mode == 3 || mode == 5);
            idesc = make_idesc_tf32_mn(64, 64, mode == 5 ? true : !a_k, mode == 5 ? true : !b_k);
            if (a_k) { da = make_kmajor_sw128_desc(smem_u32(pa)); adva = 0; }
            if (b_k) { db = make_kmajor_sw128_desc(smem_u32(pb)); advb = 0; }
        }
        umma_tf32(tm + ((uint32_t)lane_off << 16), da + adva, db + advb, idesc, 0u);
        umma_commit(&bar);
        mbar_wait(&bar, 0);
    }
    __syncthreads();
    tc_fence_after();
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < 64; c0 += 16) {
        float v[16];
        tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) D[(q * 32 + lane) * 64 + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    std::vector<float> A(ROWS * 64), B(ROWS * 64), D(128 * 64);
    for (int i = 0; i < ROWS * 64; ++i) { A[i] = (float)((i * 7 + 3) % 11 - 5); B[i] = (float)((i * 5 + 1) % 13 - 6); }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    for (int mode : {0, 1, 3, 4, 5}) for (int lane_off : {0}) for (int r8 : {0, 3}) {
        cudaMemset(dD, 0, D.size() * 4);
        k_test<<<1, 128, 100000>>>(dA, dB, dD, lane_off, r8, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0; double maxerr = 0;
        for (int m = 0; m < 64; ++m) for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < 8; ++k) ref += (double)A[(r8 * 8 + k) * 64 + m] * B[(r8 * 8 + k) * 64 + n];
            const int lane = (m % 16) + 32 * (m / 16) + lane_off;
            const double got = D[lane * 64 + n];
            if (fabs(got - ref) > 1e-3) { if (bad < 2) printf("  m=%d n=%d got %g ref %g\n", m, n, got, ref); ++bad; }
            maxerr = fmax(maxerr, fabs(got - ref));
        }
        printf("mode=%d lane_off=%d r8=%d: %d/4096 mismatches, max err %g\n", mode, lane_off, r8, bad, maxerr);
    }
    return 0;
}
