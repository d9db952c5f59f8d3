// Micro-benchmark (GPU box): scalar FFMA vs packed FFMA2 (fma.rn.f32x2) throughput on sm_100a.
// Decides whether 2-pixels-per-thread packed math is worth it for the compute-bound forward-PBR kernel.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096, ACC = 8;
__global__ void k_ffma(float* out, float a, float b) {
    float x[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0; for (int i = 0; i < ACC; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
    float2 x[ACC];
    const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.999f);
#pragma unroll
    for (int i = 0; i < ACC; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) x[i] = __ffma2_rn(x[i], A, B);
    }
    float s = 0; for (int i = 0; i < ACC; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mix(float* out, float a, float b) {   // FFMA2 + independent integer/alu work: does packing free issue slots?
    float2 x[ACC]; unsigned u[ACC];
    const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.999f);
#pragma unroll
    for (int i = 0; i < ACC; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i); u[i] = threadIdx.x + i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) { x[i] = __ffma2_rn(x[i], A, B); u[i] = (u[i] ^ 0x9e3779b9u) + (u[i] >> 3); }
    }
    float s = 0; for (int i = 0; i < ACC; ++i) s += x[i].x + x[i].y + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256;
    float* out; cudaMalloc(&out, blocks * threads * 4);
    const double n = (double)blocks * threads * ITERS * ACC;
    float t1 = timeit([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    float t2 = timeit([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    float t3 = timeit([&] { k_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    printf("SMs %d\n", p.multiProcessorCount);
    printf("FFMA : %.3f ms  %.1f TFMA/s  (%.1f TFLOP/s)\n", t1, n / t1 / 1e9, 2 * n / t1 / 1e9);
    printf("FFMA2: %.3f ms  %.1f TFMA/s  (%.1f TFLOP/s)\n", t2, 2 * n / t2 / 1e9, 4 * n / t2 / 1e9);
    printf("FFMA2+2 int ops: %.3f ms (%.1f TFMA/s)\n", t3, 2 * n / t3 / 1e9);
    return 0;
}
