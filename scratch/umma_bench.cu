// Microbenchmark: cycles per tcgen05.mma kind::tf32 (M=128, K=8) as a function of N and of how many distinct
// TMEM accumulators the back-to-back MMAs rotate over.  nvcc -gencode arch=compute_100a,code=sm_100a -I../gnn_matlang_b200/csrc
#include "tc_common.cuh"
#include <cstdio>
using namespace gnnml3;

__global__ void __launch_bounds__(128, 1) k_bench(int M, int N, int R, int iters, long long* out, int mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        // mode bit 0: A MN-major, bit 1: B MN-major (SWIZZLE_128B_BASE32B planes of 32 rows: LBO 4096, SBO 512)
        auto mn = [](uint32_t addr) {
            uint64_t d = 0;
            d |= (uint64_t)((addr & 0x3FFFF) >> 4);
            d |= (uint64_t)(4096 >> 4) << 16;
            d |= (uint64_t)(512 >> 4) << 32;
            d |= (uint64_t)1 << 46;
            d |= (uint64_t)1 << 61;
            return d;
        };
        const uint32_t idesc = make_idesc_tf32_mn(M, N, mode & 1, mode & 2);
        const uint64_t da = (mode & 1) ? mn(smem_u32(smem)) : make_kmajor_sw128_desc(smem_u32(smem));
        const uint64_t db = (mode & 2) ? mn(smem_u32(smem + 16384)) : make_kmajor_sw128_desc(smem_u32(smem + 16384));
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tm + (uint32_t)((i % R) * N);
            umma_tf32(d, da + (uint64_t)((i & 3) * ((mode & 1) ? 64 : 2)), db + (uint64_t)((i & 3) * ((mode & 2) ? 64 : 2)), idesc, i >= R ? 1u : 0u);
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

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    const int iters = 1024;
    for (int mode : {0, 1, 2, 3})
    for (int M : {64, 128})
    for (int N : {32, 64, 256}) {
        for (int R : {1, 2}) {
            if (N * R > 512) continue;
            long long h[2];
            for (int rep = 0; rep < 2; ++rep) {
                k_bench<<<1, 128, 100000>>>(M, N, R, iters, d, mode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("mode=%d M=%3d N=%3d R=%d : issue %.1f cyc/mma, complete %.1f cyc/mma\n", mode, M, N, R, (double)h[0] / iters, (double)h[1] / iters);
        }
    }
    return 0;
}
