// FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a: cycles per warp-instruction with 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, int iters, long long* cyc) {
    float a[8], b = 1.0001f, c = 0.5f;
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma2(float* out, int iters, long long* cyc) {
    unsigned long long a[8], b, c;
    float2 bb = make_float2(1.0001f, 1.0001f), cc = make_float2(0.5f, 0.5f);
    b = *reinterpret_cast<unsigned long long*>(&bb);
    c = *reinterpret_cast<unsigned long long*>(&cc);
    for (int i = 0; i < 8; ++i) { float2 t = make_float2(threadIdx.x * 0.001f + i, i); a[i] = *reinterpret_cast<unsigned long long*>(&t); }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&a[i]); s += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int warps : {4, 8, 16, 32}) {
        k_ffma<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        k_ffma<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double ffma = (double)h / (iters * 8.0);
        k_ffma2<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        k_ffma2<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double ffma2 = (double)h / (iters * 8.0);
        printf("warps/SM=%2d: FFMA %.2f cyc per warp-instr (%.1f FMA/clk/SM), FFMA2 %.2f cyc per warp-instr (%.1f FMA/clk/SM)\n", warps, ffma,
               warps * 32.0 / ffma, ffma2, warps * 64.0 / ffma2);
    }
    return 0;
}
