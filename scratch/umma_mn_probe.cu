// Brute-force probe: which shared-memory layout + descriptor (LBO, SBO, swizzle code) makes an MN-major tcgen05.mma
// kind::tf32 operand work?  One operand is MN-major (candidate layout), the other stays K-major SWIZZLE_128B (known good).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../gnn_matlang_b200/csrc umma_mn_probe.cu -o umma_mn_probe -lcuda
#include "tc_common.cuh"
#include <cstdio>
#include <vector>
using namespace gnnml3;

// element (k, m) of the MN-major operand, k < 8, m < 64 -> byte offset inside the operand's shared-memory region
__device__ __host__ inline uint32_t mn_offset(int layout, int k, int m) {
    switch (layout) {
        case 0:  // SWIZZLE_128B atoms: 32 m (128 B) x 8 k, atoms 1024 B apart along m
            return (m / 32) * 1024 + k * 128 + ((((m % 32) / 4) ^ k) << 4) + (m % 4) * 4;
        case 1:  // SWIZZLE_64B atoms: 16 m (64 B) x 8 k = 512 B
            return (m / 16) * 512 + k * 64 + ((((m % 16) / 4) ^ ((k >> 1) & 3)) << 4) + (m % 4) * 4;
        case 2:  // SWIZZLE_32B atoms: 8 m (32 B) x 8 k = 256 B
            return (m / 8) * 256 + k * 32 + ((((m % 8) / 4) ^ ((k >> 2) & 1)) << 4) + (m % 4) * 4;
        case 3:  // no swizzle: core matrices of 4 m (16 B) x 8 k = 128 B, consecutive along m
            return (m / 4) * 128 + k * 16 + (m % 4) * 4;
        case 4:  // 128-byte rows without the XOR (to see whether the swizzle is what breaks)
            return (m / 32) * 1024 + k * 128 + (m % 32) * 4;
        default: // 5: SWIZZLE_128B_BASE32B (CUTLASS Layout_MN_SW128_32B_Atom: the only MN-major layout for 32-bit operands):
                 //    atom = 4 k-rows x 128 B, 32-byte chunk index ^= k % 4; k-atoms 512 B apart (SBO), m-atoms 1024 B apart (LBO)
            return (m / 32) * 1024 + (k / 4) * 512 + (k % 4) * 128 + ((((m % 32) / 8) ^ (k % 4)) << 5) + (m % 8) * 4;
    }
}

__global__ void __launch_bounds__(128, 1) k_probe(const float* A, const float* B, float* D, int layout, uint32_t lbo, uint32_t sbo,
                                                   uint32_t swz, int which /*0: A is MN-major, 1: B is MN-major, 2: both*/) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint8_t* pa = smem;
    uint8_t* pb = smem + 8192;
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    __syncthreads();
    const bool a_mn = which == 0 || which == 2, b_mn = which == 1 || which == 2;
    for (int i = threadIdx.x; i < 8 * 64; i += blockDim.x) {
        const int k = i / 64, m = i % 64;
        const uint32_t offk = (m / 8) * 1024 + (m % 8) * 128 + (((k / 4) ^ (m % 8)) << 4) + (k % 4) * 4;
        *reinterpret_cast<float*>(pa + (a_mn ? mn_offset(layout, k, m) : offk)) = A[i];
        *reinterpret_cast<float*>(pb + (b_mn ? mn_offset(layout, k, m) : offk)) = B[i];
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 64);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        auto mk = [&](uint32_t addr) {
            uint64_t d = 0;
            d |= (uint64_t)((addr & 0x3FFFF) >> 4);
            d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
            d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
            d |= (uint64_t)1 << 46;
            d |= (uint64_t)swz << 61;
            return d;
        };
        const uint32_t idesc = make_idesc_tf32_mn(64, 64, a_mn, b_mn);
        const uint64_t da = a_mn ? mk(smem_u32(pa)) : make_kmajor_sw128_desc(smem_u32(pa));
        const uint64_t db = b_mn ? mk(smem_u32(pb)) : make_kmajor_sw128_desc(smem_u32(pb));
        umma_tf32(tm, da, db, idesc, 0u);
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
    if (threadIdx.x < 32) tmem_dealloc(tm, 64);
}

int main(int argc, char** argv) {
    const int only_which = argc > 1 ? atoi(argv[1]) : -1;
    const uint32_t only_lbo = argc > 2 ? (uint32_t)atoi(argv[2]) : 0, only_sbo = argc > 3 ? (uint32_t)atoi(argv[3]) : 0;
    std::vector<float> A(8 * 64), B(8 * 64), D(128 * 64);
    for (int i = 0; i < 8 * 64; ++i) { A[i] = (float)((i * 7 + 3) % 11 - 5); B[i] = (float)((i * 5 + 1) % 13 - 6); }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000);
    const uint32_t strides[] = {0, 16, 32, 64, 128, 256, 512, 1024, 2048};
    const int swz_of_layout[6][3] = {{2, -1, -1}, {4, -1, -1}, {6, -1, -1}, {0, -1, -1}, {0, 2, -1}, {1, -1, -1}};
    int found = 0, ran = 0, nonzero = 0;
    for (int which = 0; which < 3; ++which)
        for (int layout = 5; layout < 6; ++layout)
            for (int si = 0; si < 3 && swz_of_layout[layout][si] >= 0; ++si)
                for (uint32_t lbo : strides)
                    for (uint32_t sbo : strides) {
                        const int swz = swz_of_layout[layout][si];
                        if (only_which >= 0 && (which != only_which || lbo != only_lbo || sbo != only_sbo)) continue;
                        cudaMemset(dD, 0, D.size() * 4);
                        k_probe<<<1, 128, 20000>>>(dA, dB, dD, layout, lbo, sbo, (uint32_t)swz, which);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) {
                            printf("error %s at which=%d layout=%d swz=%d lbo=%u sbo=%u\n", cudaGetErrorString(e), which, layout, swz, lbo, sbo);
                            return 1;
                            cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
                            cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
                            cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
                            cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000);
                            continue;
                        }
                        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
                        int bad = 0, nz = 0;
                        for (int m = 0; m < 64; ++m)
                            for (int n = 0; n < 64; ++n) {
                                double ref = 0;
                                for (int k = 0; k < 8; ++k) ref += (double)A[k * 64 + m] * B[k * 64 + n];
                                const double got = D[((m % 16) + 32 * (m / 16)) * 64 + n];
                                if (fabs(got - ref) > 1e-3) ++bad;
                                if (got != 0.0) ++nz;
                            }
                        ++ran;
                        if (nz) ++nonzero;
                        if (nz && bad >= 3500 && lbo == 1024 && sbo == 1024) printf("nz   which=%d layout=%d swz=%d lbo=%u sbo=%u: %d wrong, %d nonzero\n", which, layout, swz, lbo, sbo, bad, nz);
                        if (bad == 0) { ++found; printf("OK   which=%d layout=%d swz=%d lbo=%u sbo=%u\n", which, layout, swz, lbo, sbo); }
                        else if (bad < 3500) printf("part which=%d layout=%d swz=%d lbo=%u sbo=%u: %d/4096 wrong, %d nonzero\n", which, layout, swz, lbo, sbo, bad, nz);
                    }
    printf("ran %d combinations, %d gave any nonzero output, %d exact\n", ran, nonzero, found);
    return 0;
}
