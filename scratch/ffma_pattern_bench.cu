// FP32 FMA issue rate for the aggregation pattern of the fused layer kernel: acc[k][j] += w[k] * x[j] with 4 x 16 accumulators
// per thread (three varying register operands per FFMA) versus the packed form fma.rn.f32x2 on accumulator pairs.
// 8 warps per SM (2 per scheduler), like the aggregator warps.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) k_ffma(const float* __restrict__ in, float* out, int iters, long long* cyc) {
    float acc[4][16];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[k][j] = 0.f;
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = reinterpret_cast<const float4*>(in)[i];
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const float4* p = sm + (((threadIdx.x & 31) * 5 + it * 160) % 2040);
        const float4 w4 = p[0], x0 = p[1], x1 = p[2], x2 = p[3], x3 = p[4];
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
        const float x[16] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w, x3.x, x3.y, x3.z, x3.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[k][j] = fmaf(w[k], x[j], acc[k][j]);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 16; ++j) s += acc[k][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

__global__ void __launch_bounds__(256, 1) k_ffma2(const float* __restrict__ in, float* out, int iters, long long* cyc) {
    unsigned long long acc[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k][j] = 0ull;
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = reinterpret_cast<const float4*>(in)[i];
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const float4* p = sm + (((threadIdx.x & 31) * 5 + it * 160) % 2040);
        const float4 w4 = p[0];
        const ulonglong2 x0 = *reinterpret_cast<const ulonglong2*>(p + 1), x1 = *reinterpret_cast<const ulonglong2*>(p + 2),
                         x2 = *reinterpret_cast<const ulonglong2*>(p + 3), x3 = *reinterpret_cast<const ulonglong2*>(p + 4);
        const unsigned long long w[4] = {pack2(w4.x, w4.x), pack2(w4.y, w4.y), pack2(w4.z, w4.z), pack2(w4.w, w4.w)};
        const unsigned long long x[8] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y, x3.x, x3.y};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[k][j]) : "l"(w[k]), "l"(x[j]));
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float2 t = *reinterpret_cast<float2*>(&acc[k][j]);
            s += t.x + t.y;
        }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    float *in, *out;
    long long* cyc;
    long long h;
    const int iters = 2048;
    cudaMalloc(&in, (size_t)iters * 256 * 5 * 16 + 4096);
    cudaMemset(in, 0, (size_t)iters * 256 * 5 * 16 + 4096);
    cudaMalloc(&out, 1 << 22);
    cudaMalloc(&cyc, 8);
    for (int rep = 0; rep < 2; ++rep) {
        k_ffma<<<148, 256>>>(in, out, iters, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("FFMA  : %.1f cycles per step of 64 FMA per thread, 8 warps/SM -> %.1f FMA/clk/SM (%s)\n", (double)h / iters,
               8 * 32 * 64.0 * iters / h, cudaGetErrorString(cudaGetLastError()));
        k_ffma2<<<148, 256>>>(in, out, iters, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("FFMA2 : %.1f cycles per step of 64 FMA per thread, 8 warps/SM -> %.1f FMA/clk/SM (%s)\n", (double)h / iters,
               8 * 32 * 64.0 * iters / h, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
