// Micro-benchmark (GPU box): what does a warp-wide GATHER cost in the L1 data pipe on sm_100a?
// K1's limiter after the packed-math rewrite is l1tex__data_pipe_lsu_wavefronts (73 % of peak, profiles/r02_forward_*),
// so the layout of the IBL sampling copies has to be chosen by wavefronts per footprint, not by instruction count.
// Every kernel does GATHERS loads per thread from a table much larger than L1 (L2-resident), addresses either
//   div : every lane its own pseudo-random record          coh : the 32 lanes read 32 consecutive records
// with record sizes 4/8/16/32 bytes (LDG.32/.64/.128/.256, read-only path), plus texture-path variants (tex1Dfetch float4).
// Run under ncu for the wavefront counts:  ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_tex_wavefronts.sum,
//   l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum  ./ubench_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int GATHERS = 64;
constexpr size_t TABLE_BYTES = 64ull << 20;          // 64 MB: far beyond L1, inside L2

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int BYTES> struct Rec;
template <> struct Rec<4> { using T = float; static __device__ float sum(float v) { return v; } };
template <> struct Rec<8> { using T = float2; static __device__ float sum(float2 v) { return v.x + v.y; } };
template <> struct Rec<16> { using T = float4; static __device__ float sum(float4 v) { return v.x + v.y + v.z + v.w; } };

template <int BYTES, bool COH>
__global__ void gather_kernel(const void* __restrict__ table, float* out, uint32_t nrec) {
    using T = typename Rec<BYTES>::T;
    const T* t = (const T*)table;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
#pragma unroll 8
    for (int i = 0; i < GATHERS; ++i) {
        const uint32_t h = COH ? (hash32((gid >> 5) * 977u + i) & ~31u) + (gid & 31u) : hash32(gid * 131u + i);
        acc += Rec<BYTES>::sum(__ldg(t + (h % nrec)));
    }
    out[gid] = acc;
}
template <bool COH>
__global__ void gather256_kernel(const float4* __restrict__ table, float* out, uint32_t nrec) {   // 32-byte records, LDG.E.256
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
#pragma unroll 8
    for (int i = 0; i < GATHERS; ++i) {
        const uint32_t h = COH ? (hash32((gid >> 5) * 977u + i) & ~31u) + (gid & 31u) : hash32(gid * 131u + i);
        const float4* p = table + 2u * (h % nrec);
        float a, b, c, d, e, f, g, hh;
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g), "=f"(hh) : "l"(p));
        acc += a + b + c + d + e + f + g + hh;
    }
    out[gid] = acc;
}
// two 16-byte loads of one 32-byte record (same sector) instead of one 256-bit load
__global__ void gather2x128_kernel(const float4* __restrict__ table, float* out, uint32_t nrec) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
#pragma unroll 8
    for (int i = 0; i < GATHERS; ++i) {
        const float4* p = table + 2u * (hash32(gid * 131u + i) % nrec);
        const float4 a = __ldg(p), b = __ldg(p + 1);
        acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    }
    out[gid] = acc;
}
// lane pairs cooperate: lanes 2k and 2k+1 read the two 32-byte halves of ONE 64-byte record (same 128-byte line), twice
// (once for each lane's record), and exchange halves with shuffles: 2 instructions x 16 lines instead of 2 x 32
__global__ void gather_pair64_kernel(const float4* __restrict__ table, float* out, uint32_t nrec) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    float acc = 0.0f;
#pragma unroll 4
    for (int i = 0; i < GATHERS / 2; ++i) {
        const uint32_t mine = hash32(gid * 131u + i) % nrec;                      // my 64-byte record
        const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
        for (int pass = 0; pass < 2; ++pass) {
            const uint32_t rec = ((lane & 1u) == (uint32_t)pass) ? mine : other;    // pass 0: even lane's record, pass 1: odd lane's
            const float4* p = table + 4u * rec + 2u * (lane & 1u);
            float a, b, c, d, e, f, g, hh;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g), "=f"(hh) : "l"(p));
            const float s = a + b + c + d + e + f + g + hh;
            acc += s + __shfl_xor_sync(0xffffffffu, s, 1);
        }
    }
    out[gid] = acc;
}
__global__ void gather_tex_kernel(cudaTextureObject_t tex, float* out, uint32_t nrec) {             // TEX pipe, 16-byte texels
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
#pragma unroll 8
    for (int i = 0; i < GATHERS; ++i) {
        const float4 v = tex1Dfetch<float4>(tex, (int)(hash32(gid * 131u + i) % nrec));
        acc += v.x + v.y + v.z + v.w;
    }
    out[gid] = acc;
}
// half of the gathers through the LSU pipe (LDG.128), half through the TEX pipe: do the two data paths add up?
__global__ void gather_mix_kernel(const float4* __restrict__ table, cudaTextureObject_t tex, float* out, uint32_t nrec) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
#pragma unroll 8
    for (int i = 0; i < GATHERS; i += 2) {
        const float4 a = __ldg(table + (hash32(gid * 131u + i) % nrec));
        const float4 b = tex1Dfetch<float4>(tex, (int)(hash32(gid * 131u + i + 1) % nrec));
        acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    }
    out[gid] = acc;
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 16, threads = 256;
    void* table; cudaMalloc(&table, TABLE_BYTES); cudaMemset(table, 0, TABLE_BYTES);
    float* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = table;
    rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = TABLE_BYTES;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    const double warpGathers = (double)blocks * threads / 32 * GATHERS;
    auto report = [&](const char* name, float ms, double bytesPerLane) {
        printf("%-28s %8.3f ms  %7.2f ns per warp-gather  %6.1f GB/s useful\n", name, ms, ms * 1e6 / warpGathers,
               warpGathers * 32 * bytesPerLane / ms / 1e6);
    };
    printf("SMs %d, %d warps, %d gathers per thread, table %zu MB\n", p.multiProcessorCount, blocks * threads / 32, GATHERS, TABLE_BYTES >> 20);
    report("LDG.32  divergent", timeit([&] { gather_kernel<4, false><<<blocks, threads>>>(table, out, (uint32_t)(TABLE_BYTES / 4)); }), 4);
    report("LDG.64  divergent", timeit([&] { gather_kernel<8, false><<<blocks, threads>>>(table, out, (uint32_t)(TABLE_BYTES / 8)); }), 8);
    report("LDG.128 divergent", timeit([&] { gather_kernel<16, false><<<blocks, threads>>>(table, out, (uint32_t)(TABLE_BYTES / 16)); }), 16);
    report("LDG.256 divergent", timeit([&] { gather256_kernel<false><<<blocks, threads>>>((const float4*)table, out, (uint32_t)(TABLE_BYTES / 32)); }), 32);
    report("2xLDG.128 (one 32B record)", timeit([&] { gather2x128_kernel<<<blocks, threads>>>((const float4*)table, out, (uint32_t)(TABLE_BYTES / 32)); }), 32);
    report("pair-coop 64B (2xLDG.256)", timeit([&] { gather_pair64_kernel<<<blocks, threads>>>((const float4*)table, out, (uint32_t)(TABLE_BYTES / 64)); }), 64);
    report("LDG.32  coherent", timeit([&] { gather_kernel<4, true><<<blocks, threads>>>(table, out, (uint32_t)(TABLE_BYTES / 4)); }), 4);
    report("LDG.128 coherent", timeit([&] { gather_kernel<16, true><<<blocks, threads>>>(table, out, (uint32_t)(TABLE_BYTES / 16)); }), 16);
    report("LDG.256 coherent", timeit([&] { gather256_kernel<true><<<blocks, threads>>>((const float4*)table, out, (uint32_t)(TABLE_BYTES / 32)); }), 32);
    report("TEX float4 divergent", timeit([&] { gather_tex_kernel<<<blocks, threads>>>(tex, out, (uint32_t)(TABLE_BYTES / 16)); }), 16);
    report("LDG.128 + TEX mixed", timeit([&] { gather_mix_kernel<<<blocks, threads>>>((const float4*)table, tex, out, (uint32_t)(TABLE_BYTES / 16)); }), 16);
    return 0;
}
