// microbench.cu -- measured memory-system ceilings of the B200 that bound the gather/scatter path.
// MEASURED_PEAKS.json holds only the HBM copy and cuBLAS peaks; L2 / L1 / shared-memory gather rates
// and global / shared atomic rates are measured here (SURVEY 8d asks the builder to do so).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o benchmarks/microbench benchmarks/microbench.cu
// Output: one CSV line per measurement: name,param,value,unit
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// ---- streaming read --------------------------------------------------------------------------
__global__ void stream_read(const float4* __restrict__ p, size_t n4, float* sink) {
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = p[i];
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 1234.5f) *sink = acc;
}
__global__ void stream_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

// ---- random 4-byte gathers from a global table ------------------------------------------------
template <int U>
__global__ void gather4(const float* __restrict__ t, uint32_t mask, int iters, float* sink) {
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { s = s * 1664525u + 1013904223u; v[u] = __ldg(t + (mix(s) & mask)); }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 1234.5f) *sink = acc;
}
// coherent variant: lanes of a warp read within a 32-element window (models spatially sorted points)
template <int U>
__global__ void gather4_local(const float* __restrict__ t, uint32_t mask, int iters, float* sink) {
    uint32_t s = mix((blockIdx.x * blockDim.x + threadIdx.x) / 32 + 1);
    const uint32_t lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { s = s * 1664525u + 1013904223u; v[u] = __ldg(t + (((mix(s) & ~31u) + ((lane * 7) & 31)) & mask)); }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 1234.5f) *sink = acc;
}

// ---- random shared-memory gathers -------------------------------------------------------------
template <int U>
__global__ void gather_smem(const float* __restrict__ t, int n, int iters, float* sink) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = t[i];
    __syncthreads();
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
    float acc = 0.f;
    const uint32_t mask = n - 1;
    for (int it = 0; it < iters; ++it) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { s = s * 1664525u + 1013904223u; v[u] = sm[mix(s) & mask]; }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 1234.5f) *sink = acc;
}

// ---- global float reductions (REDG.ADD.F32) -----------------------------------------------------
// group: lanes [g*group, (g+1)*group) of a warp hit the same address (1 = all distinct, 32 = whole warp)
template <int U>
__global__ void red_global(float* t, uint32_t mask, int iters, int group) {
    uint32_t s = mix((blockIdx.x * blockDim.x + threadIdx.x) / group + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = s * 1664525u + 1013904223u;
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(t + (mix(s) & mask)), "f"(1.0f) : "memory");
        }
    }
}

// ---- shared float atomicAdd (CAS loop) and shared int atomicAdd ------------------------------------
template <bool kFloat>
__global__ void atom_smem(int n, int iters, int group, float* sink) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    uint32_t s = mix((blockIdx.x * blockDim.x + threadIdx.x) / group + 1);
    const uint32_t mask = n - 1;
    for (int it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        if (kFloat) atomicAdd(&sm[mix(s) & mask], 1.0f);
        else atomicAdd(reinterpret_cast<int*>(sm) + (mix(s) & mask), 1);
    }
    __syncthreads();
    if (sm[threadIdx.x & mask] == 1234.5f) *sink = 1.f;
}

// ---- warp primitives ---------------------------------------------------------------------------------
__global__ void match_any_rate(int iters, int group, unsigned* sink) {
    unsigned key = (threadIdx.x & 31) / group, acc = 0;
    for (int it = 0; it < iters; ++it) { acc += __match_any_sync(0xffffffffu, key + (it & 1)); }
    if (acc == 12345u) *sink = acc;
}
__global__ void shfl_rate(int iters, float* sink) {
    float v = threadIdx.x, acc = 0;
    for (int it = 0; it < iters; ++it) { v = __shfl_xor_sync(0xffffffffu, v, 1 + (it & 15)); acc += v; }
    if (acc == 1234.5f) *sink = acc;
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
    printf("l2_bytes,,%d,B\n", prop.l2CacheSize);
    float* sink; CK(cudaMalloc(&sink, 64));
    const size_t big = (size_t)2 << 30;
    float *A, *B; CK(cudaMalloc(&A, big)); CK(cudaMalloc(&B, big));
    CK(cudaMemset(A, 0, big)); CK(cudaMemset(B, 0, big));

    // 1. streaming: HBM (2 GiB) and L2-resident (32 MiB re-read)
    { size_t n4 = big / 16; float ms = time_ms([&] { stream_read<<<sms * 16, 512>>>((float4*)A, n4, sink); });
      printf("stream_read_hbm,2GiB,%.1f,GB/s\n", big / ms / 1e6); }
    { size_t n4 = big / 16; float ms = time_ms([&] { stream_copy<<<sms * 16, 512>>>((float4*)A, (float4*)B, n4); });
      printf("stream_copy_hbm,2GiB,%.1f,GB/s(read+write)\n", 2.0 * big / ms / 1e6); }
    for (size_t mb : {8, 32, 64}) {
        size_t bytes = mb << 20, n4 = bytes / 16; const int rep = 50;
        stream_read<<<sms * 16, 512>>>((float4*)A, n4, sink);
        float ms = time_ms([&] { for (int r = 0; r < rep; ++r) stream_read<<<sms * 16, 512>>>((float4*)A, n4, sink); });
        printf("stream_read_l2,%zuMiB,%.1f,GB/s\n", mb, (double)bytes * rep / ms / 1e6);
    }
    // 2. random 4-byte gathers: table size sweep (L1 -> L2 -> HBM)
    for (size_t kb : {1, 16, 64, 256, 1536, 8192, 24576, 262144, 1048576}) {
        size_t entries = 1; while (entries * 4 * 2 <= kb * 1024) entries *= 2;
        const int iters = 64; const int blocks = sms * 16, threads = 256;
        float ms = time_ms([&] { gather4<16><<<blocks, threads>>>(A, (uint32_t)entries - 1, iters, sink); });
        double lanes = (double)blocks * threads * iters * 16;
        printf("gather4_random,%zuKiB,%.2f,Glanes/s\n", entries * 4 / 1024, lanes / ms / 1e6);
        ms = time_ms([&] { gather4_local<16><<<blocks, threads>>>(A, (uint32_t)entries - 1, iters, sink); });
        printf("gather4_warp_local,%zuKiB,%.2f,Glanes/s\n", entries * 4 / 1024, lanes / ms / 1e6);
    }
    // 3. shared-memory gathers
    for (int kb : {16, 64, 128}) {
        int n = kb * 256; const int iters = 256, threads = 512;
        CK(cudaFuncSetAttribute(gather_smem<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        float ms = time_ms([&] { gather_smem<16><<<sms, threads, n * 4>>>(A, n, iters, sink); });
        double lanes = (double)sms * threads * iters * 16;
        printf("gather_smem_random,%dKiB,%.2f,Glanes/s\n", kb, lanes / ms / 1e6);
    }
    // 4. global float reductions
    for (size_t kb : {1, 16, 256, 1536, 24576, 262144}) {
        size_t entries = 1; while (entries * 4 * 2 <= kb * 1024) entries *= 2;
        for (int group : {1, 4, 32}) {
            const int iters = 16, blocks = sms * 16, threads = 256;
            float ms = time_ms([&] { red_global<8><<<blocks, threads>>>(A, (uint32_t)entries - 1, iters, group); });
            double lanes = (double)blocks * threads * iters * 8;
            printf("red_global_f32,%zuKiB_group%d,%.2f,Glanes/s\n", entries * 4 / 1024, group, lanes / ms / 1e6);
        }
    }
    // 5. shared atomics
    for (int n : {64, 1024, 8192}) for (int group : {1, 4, 32}) {
        const int iters = 512, threads = 512;
        float ms = time_ms([&] { atom_smem<true><<<sms, threads, n * 4>>>(n, iters, group, sink); });
        double lanes = (double)sms * threads * iters;
        printf("atom_smem_f32_cas,%dentries_group%d,%.2f,Glanes/s\n", n, group, lanes / ms / 1e6);
        ms = time_ms([&] { atom_smem<false><<<sms, threads, n * 4>>>(n, iters, group, sink); });
        printf("atom_smem_i32,%dentries_group%d,%.2f,Glanes/s\n", n, group, lanes / ms / 1e6);
    }
    // 6. warp primitives
    for (int group : {1, 4, 32}) {
        const int iters = 4096, threads = 512;
        float ms = time_ms([&] { match_any_rate<<<sms * 4, threads>>>(iters, group, (unsigned*)sink); });
        printf("match_any,group%d,%.2f,Gwarp-instr/s\n", group, (double)sms * 4 * threads / 32 * iters / ms / 1e6);
    }
    { const int iters = 4096, threads = 512;
      float ms = time_ms([&] { shfl_rate<<<sms * 4, threads>>>(iters, sink); });
      printf("shfl_xor,,%.2f,Gwarp-instr/s\n", (double)sms * 4 * threads / 32 * iters / ms / 1e6); }
    return 0;
}
