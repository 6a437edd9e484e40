// coarse_kernels.cuh -- backward of the COARSE dense levels, accumulated in shared memory.
//
// Reference: the scatter-add of hashgrid_interpolate_cuda.cu:186-221 / hashgrid_interpolate2d_cuda.cu:199-208
// (atomicAdd(float*) straight into grad_codebook). On the coarse levels of a 3D grid millions of points add into a
// few thousand rows (level 0 of the NeRF shape: 4.2 M adds into 4913 rows), and same-address `red.global` serialises
// in the L2 slices: measured on B200, the three coarsest levels cost 40 % of the point-parallel backward.
//
// Here a dense level (or a slab of it along the slowest axis: the dense index x + y*res + z*res^2 makes a range of
// z planes a contiguous row range) lives in the 200 KB of shared memory of ONE CTA per SM. A "job" is a (level,
// slab) pair; the CTAs of a job stride over all points, keep those whose cell along the slab axis falls in the
// slab, add into shared memory and flush every touched row with one `red.global` per CTA. The point-parallel
// kernel skips the scatter of the levels handled here (`skip_mask`).
#pragma once
#include "common.cuh"

namespace shacira {

constexpr int kCoarseThreads = 512;  // half of the register file: the point-parallel kernel co-resides on the SM
constexpr int kCoarseMaxJobs = 48;
constexpr int kCoarseBudgetFloats = 50 * 1024;  // 200 KB of the 227 KB a CTA may own

struct CoarseJob {
    int32_t level;  // grid level
    int32_t c0, c1; // cells [c0, c1) along the slab axis (the last coordinate)
    int32_t row0;   // first level-local row of the slab
    int32_t rows;   // rows held in shared memory
    int32_t cta0;   // first CTA of the job
    int32_t ctas;   // CTAs of the job
    int32_t pad;
};
struct CoarseJobs {
    CoarseJob job[kCoarseMaxJobs];
    int32_t num_jobs;
    int32_t pad[3];
};

// LATENT: values per row = C, value = w_k * sum_f g[f] A[c][f] (latent_bwd_kernel); else C == F, value = w_k * g[f].
template <int D, int C, int F, bool LATENT>
__global__ void __launch_bounds__(kCoarseThreads, 2)
coarse_bwd_kernel(const float* __restrict__ coords, int64_t n, const float* __restrict__ grad_out,
                  const __grid_constant__ LevelParams lp, const __grid_constant__ CoarseJobs jobs,
                  const float* __restrict__ A, int per_level, float* __restrict__ grad_table) {
    extern __shared__ float s_acc[];
    constexpr int NV = LATENT ? C : F;
    constexpr int NC = 1 << D;
    int j = 0;
    while (j + 1 < jobs.num_jobs && (int)blockIdx.x >= jobs.job[j + 1].cta0) ++j;
    const CoarseJob jb = jobs.job[j];
    const int l = jb.level, L = lp.num_lods;
    const int nvals = jb.rows * NV;
    for (int e = threadIdx.x; e < nvals; e += kCoarseThreads) s_acc[e] = 0.0f;
    float a[C][F];
    if constexpr (LATENT) {
        const int la = per_level ? l : 0;
#pragma unroll
        for (int ch = 0; ch < C; ++ch)
#pragma unroll
            for (int f = 0; f < F; ++f) a[ch][f] = __ldg(A + (la * C + ch) * F + f);
    }
    __syncthreads();
    const int32_t res = lp.res[l];
    const float hi = lp.hi[l];
    const bool vec_g = (F == 1) || ((L * F) % (F >= 4 ? 4 : F) == 0);
    const int part = (int)blockIdx.x - jb.cta0;
    for (int64_t i = (int64_t)part * kCoarseThreads + threadIdx.x; i < n; i += (int64_t)jb.ctas * kCoarseThreads) {
        double t[D];
        load_unit_coords<D>(coords, i, t);
        int32_t pc;
        float fc, gc;
        locate(t[D - 1], res, hi, pc, fc, gc);
        if (pc < jb.c0 || pc >= jb.c1) continue;
        Corners<D> c;
        corners<D>(t, lp, l, c);
        float g[F];
        const float* g_row = grad_out + i * (int64_t)L * F + l * F;
        if (vec_g) {
            load_row<F>(g_row, g);
        } else {
#pragma unroll
            for (int f = 0; f < F; ++f) g[f] = __ldg(g_row + f);
        }
        float gz[NV];
        if constexpr (LATENT) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                float acc = 0.0f;
#pragma unroll
                for (int f = 0; f < F; ++f) acc = __fmaf_rn(g[f], a[ch][f], acc);
                gz[ch] = acc;
            }
        } else {
#pragma unroll
            for (int f = 0; f < F; ++f) gz[f] = g[f];
        }
        // shared-memory float adds compile to compare-and-swap loops (ATOMS.CAST.SPIN; no native ATOMS.ADD.F32).
        // Measured on B200: software-pipelined loads and interleaving the corners' loops by hand made this kernel
        // 1.6x SLOWER -- it is bound by the CAS rate of the shared-memory pipe (~0.75 lanes/clk/SM), not by latency.
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int32_t r = c.idx[k] - jb.row0;
            if ((uint32_t)r < (uint32_t)jb.rows) {
#pragma unroll
                for (int ch = 0; ch < NV; ++ch) atomicAdd(&s_acc[r * NV + ch], __fmul_rn(gz[ch], c.w[k]));
            } else {  // a corner outside the slab (only the zero-weight corners of SURVEY Q4 can get here)
#pragma unroll
                for (int ch = 0; ch < NV; ++ch)
                    red_add(grad_table + ((int64_t)lp.first[l] + c.idx[k]) * NV + ch, __fmul_rn(gz[ch], c.w[k]));
            }
        }
    }
    __syncthreads();
    // flush: the CTAs of a job start at different offsets so that they do not walk the same rows in lock step
    float* base = grad_table + ((int64_t)lp.first[l] + jb.row0) * NV;
    const int start = (int)(((int64_t)part * nvals) / jb.ctas);
    for (int e = threadIdx.x; e < nvals; e += kCoarseThreads) {
        int ee = e + start;
        if (ee >= nvals) ee -= nvals;
        const float v = s_acc[ee];
        if (v != 0.0f) red_add(base + ee, v);
    }
}

// Host side: which dense levels go to shared memory, cut into slabs, and how the CTAs are shared out.
// Returns the mask of the levels covered (0: nothing to do). `max_slabs` bounds the scan overhead per level.
inline uint32_t plan_coarse_jobs(int dim, const LevelParams& lp, int nv, int64_t n, int sms, int max_slabs,
                                 CoarseJobs& jobs, uint32_t level_mask = 0xffffffffu) {
    memset(&jobs, 0, sizeof(jobs));
    uint32_t mask = 0;
    double weight[kCoarseMaxJobs];
    double wsum = 0.0;
    for (int l = 0; l < lp.num_lods; ++l) {
        if (!((lp.dense_mask >> l) & 1u) || !((level_mask >> l) & 1u)) continue;
        const int64_t res = lp.res[l];
        const int64_t plane = dim == 2 ? res : res * res;
        const int64_t max_planes = kCoarseBudgetFloats / (nv * plane);
        if (max_planes < 2) continue;
        const int64_t h = max_planes - 1;                 // cells per slab (a slab of h cells touches h + 1 planes)
        const int64_t slabs = (res + h - 1) / h;
        if (slabs > max_slabs || jobs.num_jobs + slabs > kCoarseMaxJobs) continue;
        // adds per row of the level: below ~8 the direct scatter is not contended and privatising only adds a flush
        if ((double)n * (1 << dim) / (double)lp.rows[l] < 8.0) continue;
        for (int64_t s = 0; s < slabs; ++s) {
            CoarseJob& jb = jobs.job[jobs.num_jobs];
            jb.level = l;
            jb.c0 = (int32_t)(s * h);
            jb.c1 = (int32_t)((s + 1 == slabs) ? res : (s + 1) * h);  // the last slab takes every remaining cell
            jb.row0 = (int32_t)(jb.c0 * plane);
            int64_t rows = (int64_t)(jb.c1 - jb.c0 + 1) * plane;
            if (jb.row0 + rows > lp.rows[l]) rows = lp.rows[l] - jb.row0;
            jb.rows = (int32_t)rows;
            // cost model: every point is scanned, the points of the slab make 2^dim shared-memory adds
            weight[jobs.num_jobs] = 1.0 + 2.0 * (1 << dim) * (double)(jb.c1 - jb.c0) / (double)res;
            wsum += weight[jobs.num_jobs];
            ++jobs.num_jobs;
        }
        mask |= (1u << l);
    }
    if (!jobs.num_jobs) return 0;
    int left = sms - jobs.num_jobs;  // one CTA each, the rest by weight
    if (left < 0) left = 0;
    int cta = 0;
    for (int j = 0; j < jobs.num_jobs; ++j) {
        int extra = (int)(left * weight[j] / wsum);
        jobs.job[j].cta0 = cta;
        jobs.job[j].ctas = 1 + extra;
        cta += jobs.job[j].ctas;
    }
    return mask;
}

inline int coarse_total_ctas(const CoarseJobs& jobs) {
    const CoarseJob& last = jobs.job[jobs.num_jobs - 1];
    return last.cta0 + last.ctas;
}

inline size_t coarse_smem_bytes(const CoarseJobs& jobs, int nv) {
    int32_t rows = 0;
    for (int j = 0; j < jobs.num_jobs; ++j) rows = jobs.job[j].rows > rows ? jobs.job[j].rows : rows;
    return sizeof(float) * (size_t)rows * nv;
}

}  // namespace shacira
