// peer_kernels.cuh -- the exchange step of the NeRF ray-batch data-parallel path over NVLink / NVSwitch peer memory.
//
// BASELINE.json north_star: "NeRF ray batches are data-parallel, with the hash-table/latent gradient allreduced ... over
// NVLink". The reference itself is single-GPU (no distributed code). The NCCL form of this step --
// one ncclAllReduce of the 24.4 MB gradient arena -- leaves ~125 us exposed at 8 GPUs (profiles/r02k_bench_n8.json).
//
// Here every rank maps every other rank's gradient arena (CUDA IPC, or plain peer access inside one process) and ONE
// kernel per rank does the whole exchange:
//     barrier A  (every rank's producers have finished: flag written into each peer, local spin)
//     rank r owns slice r of the arena: for each 16-byte piece, load it from all N arenas (N independent NVLink reads
//     in flight per thread), add in rank order 0..N-1, store the sum into all N arenas
//     barrier B  (the last CTA of a rank tells every peer "my stores are out"; a rank's kernel ends when it has heard
//     that from every peer)
// Each element is reduced on exactly one rank, in a fixed order, and the same value is written everywhere: all ranks end
// with BIT-IDENTICAL sums (replicated parameters cannot drift apart), independent of timing. Per GPU (N-1)/N of the
// arena travels in and out once: 21 MB each way at N = 8 against NVLink 5's 900 GB/s per direction.
// With `adam` set the owner applies the table's Adam update to its slice instead and broadcasts the updated PARAMETERS
// (reduce-scatter + sharded optimizer state + all-gather in the same pass).
#pragma once
#include "common.cuh"

namespace shacira {

constexpr int kPeerMax = 8;
constexpr int kPeerThreads = 512;

struct PeerView {
    float* buf[kPeerMax];          // every rank's arena (index = rank), as mapped into THIS process
    unsigned* flags[kPeerMax];     // every rank's flag block: [0, 8) barrier A, [8, 16) barrier B, [16] epoch, [17] ticket, [18] error word (barrier timeout)
    int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer(const float* p) {   // coherent at system scope, not cached in L1
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer(float* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Spin until *p has reached `epoch` (wrap-safe). Bounded: ~4 s of polling, then the kernel gives up, raises the error word
// of the local flag block (slot 18) and goes on -- a peer that died must not hang this GPU forever (the host reads the
// word with shacira_peer_status; the data of such a call is garbage).
__device__ __forceinline__ void peer_wait(const unsigned* p, unsigned epoch, unsigned* err) {
    for (unsigned spins = 0; (int)(ld_acquire_sys(p) - epoch) < 0; ++spins) {
        if (spins > 64u) __nanosleep(200);
        if (spins > 20000000u) { atomicExch(err, 1u); return; }
    }
}

struct PeerAdam {       // optional fused optimizer on the owner's slice (all pointers in the owner's memory, nullptr = off)
    float* param[kPeerMax];   // every rank's parameter table (the update is broadcast)
    float* exp_avg;           // this rank's slice-local state, indexed like the arena
    float* exp_avg_sq;
    const float* step;        // device float counter (number of steps taken so far)
    float lr, beta1, beta2, eps, weight_decay;
    int64_t numel;            // leading elements of the arena that are the table's gradient
};

template <int N, bool ADAM>
__global__ void __launch_bounds__(kPeerThreads)
peer_allreduce_kernel(const __grid_constant__ PeerView pv, int64_t numel4, const __grid_constant__ PeerAdam ad) {
    unsigned* mine = pv.flags[pv.rank];
    const unsigned epoch = ld_acquire_sys(mine + 16) + 1u;   // same on every rank: calls are made in lockstep
    // ---- barrier A ---------------------------------------------------------------------------------------------------
    if (blockIdx.x == 0 && threadIdx.x < N) {
        __threadfence_system();
        st_release_sys(pv.flags[threadIdx.x] + pv.rank, epoch);       // "rank pv.rank has arrived" into peer threadIdx.x
    }
    if (threadIdx.x < N) {
        peer_wait(mine + threadIdx.x, epoch, mine + 18);   // local spin: peers write into my flags
    }
    __syncthreads();
    // ---- slice pv.rank: reduce from all arenas, broadcast into all arenas -----------------------------------------------
    const int64_t per = (numel4 + N - 1) / N;
    const int64_t begin = per * pv.rank, end = min(numel4, begin + per);
    float bc1 = 1.0f, bc2s = 1.0f;
    if (ADAM) {
        const float tstep = *ad.step + 1.0f;
        bc1 = 1.0f - powf(ad.beta1, tstep);
        bc2s = rsqrtf(1.0f - powf(ad.beta2, tstep));
    }
    for (int64_t i = begin + (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; i < end; i += (int64_t)gridDim.x * kPeerThreads) {
        float4 v[N];
#pragma unroll
        for (int p = 0; p < N; ++p) v[p] = ld_peer(pv.buf[p] + 4 * i);
        float4 s = v[0];
#pragma unroll
        for (int p = 1; p < N; ++p) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
        if (ADAM && 4 * i + 3 < ad.numel) {
            // torch.optim.Adam on the owner's slice; every rank receives the updated parameters, the gradient slots
            // are cleared for the next step
            float4 pp = *reinterpret_cast<const float4*>(ad.param[pv.rank] + 4 * i);
            float4 mm = *reinterpret_cast<float4*>(ad.exp_avg + 4 * i), vv = *reinterpret_cast<float4*>(ad.exp_avg_sq + 4 * i);
            float* P = &pp.x; float* M = &mm.x; float* V = &vv.x; const float* G = &s.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gk = fmaf(ad.weight_decay, P[k], G[k]);
                M[k] = fmaf(ad.beta1, M[k], (1.0f - ad.beta1) * gk);
                V[k] = fmaf(ad.beta2, V[k], (1.0f - ad.beta2) * gk * gk);
                P[k] -= (ad.lr / bc1) * M[k] / (sqrtf(V[k]) * bc2s + ad.eps);
            }
            *reinterpret_cast<float4*>(ad.exp_avg + 4 * i) = mm;
            *reinterpret_cast<float4*>(ad.exp_avg_sq + 4 * i) = vv;
#pragma unroll
            for (int p = 0; p < N; ++p) st_peer(ad.param[p] + 4 * i, pp);
#pragma unroll
            for (int p = 0; p < N; ++p) st_peer(pv.buf[p] + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
        } else {
#pragma unroll
            for (int p = 0; p < N; ++p) st_peer(pv.buf[p] + 4 * i, s);
        }
    }
    // ---- barrier B: the last CTA of this rank announces "my stores are out" and waits for everybody else's ----------------
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_last = (atomicAdd(mine + 17, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x < N) {
        __threadfence_system();
        st_release_sys(pv.flags[threadIdx.x] + 8 + pv.rank, epoch);
        peer_wait(mine + 8 + threadIdx.x, epoch, mine + 18);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mine[17] = 0u;            // ticket ready for the next call
        st_release_sys(mine + 16, epoch);
    }
}

// ---- the same exchange through the NVSwitch multicast object (NVLS) ---------------------------------------------------------
// `mc` is the multicast address of the arena (every rank's copy bound to one multicast object; the host side maps it --
// shacira_b200/peer.py uses torch's symmetric-memory allocator for that plumbing). multimem.ld_reduce returns the SUM of
// the N copies, added inside the switch; multimem.st stores to all N copies. Per GPU and direction one arena's worth of
// bytes crosses NVLink instead of 2 (N-1)/N arenas; the barriers are the ones above (flags in plain peer memory).
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int N>
__global__ void __launch_bounds__(kPeerThreads)
peer_allreduce_mc_kernel(const __grid_constant__ PeerView pv, float* __restrict__ mc, int64_t numel4) {
    unsigned* mine = pv.flags[pv.rank];
    const unsigned epoch = ld_acquire_sys(mine + 16) + 1u;
    if (blockIdx.x == 0 && threadIdx.x < N) {
        __threadfence_system();
        st_release_sys(pv.flags[threadIdx.x] + pv.rank, epoch);
    }
    if (threadIdx.x < N) {
        peer_wait(mine + threadIdx.x, epoch, mine + 18);
    }
    __syncthreads();
    const int64_t per = (numel4 + N - 1) / N;
    const int64_t begin = per * pv.rank, end = min(numel4, begin + per);
    constexpr int U = 4;   // pieces in flight per thread
    const int64_t stride = (int64_t)gridDim.x * kPeerThreads;
    for (int64_t i0 = begin + (int64_t)blockIdx.x * kPeerThreads + threadIdx.x; i0 < end; i0 += stride * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i0 + u * stride < end) v[u] = multimem_ld_reduce_add(mc + 4 * (i0 + u * stride));
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i0 + u * stride < end) multimem_st(mc + 4 * (i0 + u * stride), v[u]);
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_last = (atomicAdd(mine + 17, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x < N) {
        __threadfence_system();
        st_release_sys(pv.flags[threadIdx.x] + 8 + pv.rank, epoch);
        peer_wait(mine + 8 + threadIdx.x, epoch, mine + 18);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mine[17] = 0u;
        st_release_sys(mine + 16, epoch);
    }
}

}  // namespace shacira
