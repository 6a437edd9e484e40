// microbench2.cu -- round-2 design questions of the 3D path, measured on the B200:
//   (a) the two x-neighbour corners of a cell almost always share a 32-byte sector of the table: what does it cost to
//       fetch them with two scalar loads (what the point-parallel kernel does) against ONE 8 / 16-byte load?
//   (b) the same for the backward's float reductions: two scalar `red.global.add.f32` against one `.v2` / `.v4`.
//   (c) warp-local (sorted) variants of both.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o benchmarks/microbench2 benchmarks/microbench2.cu
// Output: one CSV line per measurement: name,param,value,unit   (value = corner PAIRS per second)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// MODE 0: two scalar loads (a, a^1) issued U apart like corners k and k+4 of the point-parallel kernel
// MODE 1: one 8-byte load of the aligned pair
// MODE 2: one 16-byte load of the aligned quad
// MODE 3: ONE scalar load (reference: half the lanes)
template <int U, int MODE>
__global__ void gather_pairs(const float* __restrict__ t, uint32_t mask, int iters, float* sink) {
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        uint32_t idx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { s = s * 1664525u + 1013904223u; idx[u] = mix(s) & mask; }
        if (MODE == 0) {
            float v[2 * U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldg(t + idx[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) v[U + u] = __ldg(t + (idx[u] ^ 1u));
#pragma unroll
            for (int u = 0; u < 2 * U; ++u) acc += v[u];
        } else if (MODE == 1) {
            float2 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const float2*>(t) + (idx[u] >> 1));
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y;
        } else if (MODE == 2) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(t) + (idx[u] >> 2));
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        } else {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldg(t + idx[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u];
        }
    }
    if (acc == 1234.5f) *sink = acc;
}

// MODE 0: two scalar red (a, a^1); 1: one red.v2 on the aligned pair; 2: one red.v4 on the aligned quad (two zero lanes);
// 3: one scalar red
template <int U, int MODE>
__global__ void red_pairs(float* t, uint32_t mask, int iters) {
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t i = mix(s) & mask;
            if (MODE == 0) {
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(t + i), "f"(1.0f) : "memory");
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(t + (i ^ 1u)), "f"(1.0f) : "memory");
            } else if (MODE == 1) {
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(t + (i & ~1u)), "f"(1.0f), "f"(1.0f) : "memory");
            } else if (MODE == 2) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(t + (i & ~3u)), "f"(1.0f), "f"(1.0f), "f"(0.0f), "f"(0.0f) : "memory");
            } else {
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(t + i), "f"(1.0f) : "memory");
            }
        }
    }
}

// lane pairs: lanes 2i and 2i+1 of a warp access the two rows of ONE x-pair (a, a ^ k) in the same instruction
// (k = 1: same 8 bytes; k = 7: same sector; k = 31: same 128-byte line)
template <int U>
__global__ void gather_lanepair(const float* __restrict__ t, uint32_t mask, int iters, uint32_t k, float* sink) {
    uint32_t s = mix((blockIdx.x * blockDim.x + threadIdx.x) / 2 + 1);
    const uint32_t odd = threadIdx.x & 1;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { s = s * 1664525u + 1013904223u; const uint32_t a = mix(s) & mask; v[u] = __ldg(t + (odd ? (a ^ k) : a)); }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 1234.5f) *sink = acc;
}
template <int U>
__global__ void red_lanepair(float* t, uint32_t mask, int iters, uint32_t k) {
    uint32_t s = mix((blockIdx.x * blockDim.x + threadIdx.x) / 2 + 1);
    const uint32_t odd = threadIdx.x & 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t a = mix(s) & mask;
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(t + (odd ? (a ^ k) : a)), "f"(1.0f) : "memory");
        }
    }
}

template <typename F>
float time_ms(F launch, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device,%s,%d,SMs\n", prop.name, sms);
    float* sink; CK(cudaMalloc(&sink, 64));
    const size_t big = (size_t)256 << 20;
    float* A; CK(cudaMalloc(&A, big));
    CK(cudaMemset(A, 0, big));
    const int blocks = sms * 16, threads = 256;
    for (size_t kb : {2048, 24576}) {   // one hashed level (2^19 rows) / the whole cfg4 table
        size_t entries = 1; while (entries * 4 * 2 <= kb * 1024) entries *= 2;
        const uint32_t mask = (uint32_t)entries - 1;
        const int iters = 32;
        const double pairs = (double)blocks * threads * iters * 8;
        float ms;
        ms = time_ms([&] { gather_pairs<8, 0><<<blocks, threads>>>(A, mask, iters, sink); });
        printf("gather_pair_2xLDG32,%zuKiB,%.2f,Gpairs/s\n", entries * 4 / 1024, pairs / ms / 1e6);
        ms = time_ms([&] { gather_pairs<8, 1><<<blocks, threads>>>(A, mask, iters, sink); });
        printf("gather_pair_LDG64,%zuKiB,%.2f,Gpairs/s\n", entries * 4 / 1024, pairs / ms / 1e6);
        ms = time_ms([&] { gather_pairs<8, 2><<<blocks, threads>>>(A, mask, iters, sink); });
        printf("gather_pair_LDG128,%zuKiB,%.2f,Gpairs/s\n", entries * 4 / 1024, pairs / ms / 1e6);
        ms = time_ms([&] { gather_pairs<8, 3><<<blocks, threads>>>(A, mask, iters, sink); });
        printf("gather_single_LDG32,%zuKiB,%.2f,Glanes/s\n", entries * 4 / 1024, pairs / ms / 1e6);
        const int ri = 16;
        const double rp = (double)blocks * threads * ri * 8;
        ms = time_ms([&] { red_pairs<8, 0><<<blocks, threads>>>(A, mask, ri); });
        printf("red_pair_2xRED32,%zuKiB,%.2f,Gpairs/s\n", entries * 4 / 1024, rp / ms / 1e6);
        ms = time_ms([&] { red_pairs<8, 1><<<blocks, threads>>>(A, mask, ri); });
        printf("red_pair_REDv2,%zuKiB,%.2f,Gpairs/s\n", entries * 4 / 1024, rp / ms / 1e6);
        ms = time_ms([&] { red_pairs<8, 2><<<blocks, threads>>>(A, mask, ri); });
        printf("red_pair_REDv4,%zuKiB,%.2f,Gpairs/s\n", entries * 4 / 1024, rp / ms / 1e6);
        ms = time_ms([&] { red_pairs<8, 3><<<blocks, threads>>>(A, mask, ri); });
        printf("red_single_RED32,%zuKiB,%.2f,Glanes/s\n", entries * 4 / 1024, rp / ms / 1e6);
        for (uint32_t k : {1u, 7u, 31u, 1023u}) {
            ms = time_ms([&] { gather_lanepair<8><<<blocks, threads>>>(A, mask, iters, k, sink); });
            printf("gather_lanepair_xor%u,%zuKiB,%.2f,Gpairs/s\n", k, entries * 4 / 1024, pairs / 2 / ms / 1e6);
            ms = time_ms([&] { red_lanepair<8><<<blocks, threads>>>(A, mask, ri, k); });
            printf("red_lanepair_xor%u,%zuKiB,%.2f,Gpairs/s\n", k, entries * 4 / 1024, rp / 2 / ms / 1e6);
        }
    }
    return 0;
}
