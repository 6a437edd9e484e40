/*
 * hashgrid_oracle.c -- CPU restatement of the reference's multi-level hash-grid
 * interpolate kernels. TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA path in shacira_b200/csrc. It may be
 * imported / linked / executed only by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / `--impl reference` legs of bench.py. The product path never calls it.
 *
 * It follows the reference line by line (paths relative to /root/reference):
 *   hash_index2d            wisp/csrc/ops/hashgrid_interpolate2d_cuda.cu:17-36
 *   2D forward              wisp/csrc/ops/hashgrid_interpolate2d_cuda.cu:44-99
 *   2D backward             wisp/csrc/ops/hashgrid_interpolate2d_cuda.cu:133-208
 *   hash_index (3D)         wisp/csrc/ops/hashgrid_interpolate_cuda.cu:17-39
 *   3D forward              wisp/csrc/ops/hashgrid_interpolate_cuda.cu:47-109
 *   3D backward             wisp/csrc/ops/hashgrid_interpolate_cuda.cu:143-221
 *   host level loop         wisp/csrc/ops/hashgrid_interpolate.cpp:44-100,130-186
 *
 * Arithmetic that is restated exactly:
 *   - x = resolution * (coord * 0.5 + 0.5) is evaluated in DOUBLE and narrowed to float
 *     when passed to clamp(float,float,float) (2d_cuda.cu:65-66, _cuda.cu:68-70);
 *   - the clamp bounds are (float)0 and (float)(resolution - 1 - 1e-5) (double -> float);
 *   - pos = floor(x) (float floor, then int conversion); frac = x - (float)pos;
 *     1 - frac is a double subtraction narrowed to float (2d_cuda.cu:71);
 *   - the dense-vs-hash predicate and the dense index use int32 arithmetic
 *     (wrap-around included), the hash uses uint32 wrap-around products,
 *     xor and `% codebook_size`;
 *   - corner j: 2D x += (j>>1)&1, y += j&1; 3D x += (j>>2)&1, y += (j>>1)&1, z += j&1;
 *   - features: sum over corners k = 0..2^D-1 in that order. The reference is compiled
 *     with nvcc's default -fmad=true, which contracts v0*c0 + v1*c1 + v2*c2 + ... into
 *     fma(v3,c3, fma(v2,c2, fma(v0,c0, v1*c1))) -- the first sum a*b + c*d becomes
 *     fma(a,b, c*d), as read from the sm_100a SASS of the reference's float kernel
 *     (FMUL by c001, then FFMA by c000, c010, c011). ORACLE_FMA selects that form
 *     (default) or separately rounded mul/add. The two differ by <= 1 ulp per term,
 *     well inside the 1e-5 feature tolerance.
 *   - backward: grad_codebook[idx_k*F + j] += grad_out[i, lod*F + j] * c_k.
 *     The reference uses float atomicAdd (order non-deterministic); here each thread
 *     accumulates a private copy and copies are summed in thread order (deterministic).
 *
 * Parity pin: the reference ships no CPU implementation, tests or golden vectors for
 * this path. The pin is (1) oracle/_ref -- the reference's own .cu
 * files compiled unmodified for sm_100a (oracle/build_ref.py) and compared with this file
 * on the GPU box (tests/test_ref_kernels_gpu.py: forward bit-identical), and (2)
 * tests/golden/hashgrid_ref_kernels.npz, outputs of those reference kernels captured on a
 * B200 by tests/golden/make_golden_gpu.py and checked on CPU (tests/test_oracle_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef ORACLE_FMA
#define ORACLE_FMA 1
#endif

static inline float clampf(float x, float a, float b) { return fmaxf(a, fminf(b, x)); }

/* hashgrid_interpolate2d_cuda.cu:17-36 */
static inline int32_t hash_index2d(int32_t px, int32_t py, int32_t resolution, int32_t codebook_size) {
    /* int32 products wrap like the device code (signed overflow is UB in C, so do it unsigned) */
    int32_t rr = (int32_t)((uint32_t)resolution * (uint32_t)resolution);
    if (resolution < codebook_size && rr < codebook_size) {
        return (int32_t)((uint32_t)px + (uint32_t)py * (uint32_t)resolution);
    }
    uint32_t h = ((uint32_t)px * 1u) ^ ((uint32_t)py * 2654435761u);
    return (int32_t)(h % (uint32_t)codebook_size);
}

/* hashgrid_interpolate_cuda.cu:17-39 */
static inline int32_t hash_index3d(int32_t px, int32_t py, int32_t pz, int32_t resolution, int32_t codebook_size) {
    int32_t rr = (int32_t)((uint32_t)resolution * (uint32_t)resolution);
    int32_t rrr = (int32_t)((uint32_t)rr * (uint32_t)resolution);
    if (resolution < codebook_size && rr < codebook_size && rrr < codebook_size) {
        return (int32_t)((uint32_t)px + (uint32_t)py * (uint32_t)resolution +
                         (uint32_t)pz * (uint32_t)resolution * (uint32_t)resolution);
    }
    uint32_t h = ((uint32_t)px * 1u) ^ ((uint32_t)py * 2654435761u) ^ ((uint32_t)pz * 805459861u);
    return (int32_t)(h % (uint32_t)codebook_size);
}

/* Coordinate -> (cell, fractional offset). 2d_cuda.cu:65-71 / _cuda.cu:68-76. */
static inline void cell_of(float coord, int32_t resolution, int32_t* pos, float* frac, float* one_minus) {
    float hi = (float)((double)(resolution - 1) - 1e-5);
    float x = clampf((float)((double)resolution * ((double)coord * 0.5 + 0.5)), 0.0f, hi);
    int32_t p = (int32_t)floorf(x);
    float f = x - (float)p;
    *pos = p;
    *frac = f;
    *one_minus = (float)(1.0 - (double)f);
}

static inline float corner_sum(const float* v, const float* c, int n) {
#if ORACLE_FMA
    float acc = v[1] * c[1];
    acc = fmaf(v[0], c[0], acc);
    for (int k = 2; k < n; ++k) acc = fmaf(v[k], c[k], acc);
    return acc;
#else
    volatile float acc = v[0] * c[0];
    for (int k = 1; k < n; ++k) { volatile float p = v[k] * c[k]; acc = acc + p; }
    return acc;
#endif
}

/* Corner indices and weights of one point at one level. Returns 2^dim. */
static inline int corners_of(int dim, const float* coord, int32_t resolution, int32_t codebook_size,
                             int32_t* idx, float* w) {
    int32_t p[3];
    float f[3], g[3];
    for (int d = 0; d < dim; ++d) cell_of(coord[d], resolution, &p[d], &f[d], &g[d]);
    if (dim == 2) {
        /* 2d_cuda.cu:72-75 */
        w[0] = g[0] * g[1];
        w[1] = g[0] * f[1];
        w[2] = f[0] * g[1];
        w[3] = f[0] * f[1];
        for (int j = 0; j < 4; ++j)
            idx[j] = hash_index2d(p[0] + ((j & 2) >> 1), p[1] + (j & 1), resolution, codebook_size);
        return 4;
    }
    /* _cuda.cu:77-84: (a*b)*c, left to right */
    w[0] = g[0] * g[1] * g[2];
    w[1] = g[0] * g[1] * f[2];
    w[2] = g[0] * f[1] * g[2];
    w[3] = g[0] * f[1] * f[2];
    w[4] = f[0] * g[1] * g[2];
    w[5] = f[0] * g[1] * f[2];
    w[6] = f[0] * f[1] * g[2];
    w[7] = f[0] * f[1] * f[2];
    for (int j = 0; j < 8; ++j)
        idx[j] = hash_index3d(p[0] + ((j & 4) >> 2), p[1] + ((j & 2) >> 1), p[2] + (j & 1), resolution,
                              codebook_size);
    return 8;
}

/*
 * Corner indices (level-local, as the reference computes them) and weights for every
 * point/level: idx_out[N, L, 2^dim] int32, w_out[N, L, 2^dim] float. Used by the
 * bit-exactness tests for hash indices.
 */
void oracle_hashgrid_corners(int dim, const float* coords, int64_t n, const int32_t* resolutions, int32_t num_lods,
                             int32_t codebook_bitwidth, int32_t* idx_out, float* w_out) {
    const int32_t codebook_size = (int32_t)pow(2, codebook_bitwidth); /* hashgrid_interpolate.cpp:56 */
    const int nc = 1 << dim;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        for (int32_t l = 0; l < num_lods; ++l) {
            corners_of(dim, coords + i * dim, resolutions[l], codebook_size, idx_out + (i * num_lods + l) * nc,
                       w_out + (i * num_lods + l) * nc);
        }
    }
}

/*
 * Forward. feats[N, L*F], element [i, lod*F + j] (2d_cuda.cu:96). `table_entries` bounds
 * the reads: the reference reads out of range with weight exactly 0 on dense levels with
 * res >= 257; such reads are redirected to entry 0 of the level here (the
 * product 0 * finite is 0 either way). Returns the number of redirected reads.
 */
int64_t oracle_hashgrid_forward(int dim, const float* coords, int64_t n, const float* codebook, int64_t table_entries,
                                const int32_t* first_idx, const int32_t* resolutions, int32_t num_lods,
                                int32_t codebook_bitwidth, int32_t feature_dim, float* feats) {
    const int32_t codebook_size = (int32_t)pow(2, codebook_bitwidth);
    int64_t redirected = 0;
#pragma omp parallel for schedule(static) reduction(+ : redirected)
    for (int64_t i = 0; i < n; ++i) {
        int32_t idx[8];
        float w[8], v[8];
        for (int32_t l = 0; l < num_lods; ++l) {
            const int nc = corners_of(dim, coords + i * dim, resolutions[l], codebook_size, idx, w);
            const float* cb = codebook + (int64_t)first_idx[l] * feature_dim;
            const int64_t room = table_entries - first_idx[l];
            for (int k = 0; k < nc; ++k) {
                if (idx[k] < 0 || idx[k] >= room) { idx[k] = 0; ++redirected; }
            }
            for (int32_t j = 0; j < feature_dim; ++j) {
                for (int k = 0; k < nc; ++k) v[k] = cb[(int64_t)idx[k] * feature_dim + j];
                feats[(int64_t)num_lods * i * feature_dim + (int64_t)feature_dim * l + j] = corner_sum(v, w, nc);
            }
        }
    }
    return redirected;
}

/*
 * Backward. grad_codebook[T, F] is overwritten (the reference allocates zeros_like,
 * hashgrid_interpolate.cpp:81,167). Level-outer, thread-private accumulation.
 */
int64_t oracle_hashgrid_backward(int dim, const float* coords, int64_t n, const float* grad_output,
                                 int64_t table_entries, const int32_t* first_idx, const int32_t* resolutions,
                                 int32_t num_lods, int32_t codebook_bitwidth, int32_t feature_dim,
                                 float* grad_codebook) {
    const int32_t codebook_size = (int32_t)pow(2, codebook_bitwidth);
    int64_t redirected = 0;
    memset(grad_codebook, 0, sizeof(float) * (size_t)table_entries * feature_dim);
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    for (int32_t l = 0; l < num_lods; ++l) {
        const int64_t lvl_begin = first_idx[l];
        const int64_t lvl_end = (l + 1 < num_lods) ? first_idx[l + 1] : table_entries;
        /* reads past the level (Q4) carry weight 0; keep one slack row per level for them */
        const int64_t span = (lvl_end - lvl_begin) * feature_dim;
        float* priv = (float*)calloc((size_t)span * nthreads, sizeof(float));
#pragma omp parallel reduction(+ : redirected)
        {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            float* mine = priv + (size_t)span * tid;
            int32_t idx[8];
            float w[8];
#pragma omp for schedule(static)
            for (int64_t i = 0; i < n; ++i) {
                const int nc = corners_of(dim, coords + i * dim, resolutions[l], codebook_size, idx, w);
                for (int32_t j = 0; j < feature_dim; ++j) {
                    const float g = grad_output[i * num_lods * feature_dim + (int64_t)l * feature_dim + j];
                    for (int k = 0; k < nc; ++k) {
                        int64_t e = idx[k];
                        if (e < 0 || e >= lvl_end - lvl_begin) {
                            /* out-of-level target: weight is exactly 0 in the reference (Q4) */
                            ++redirected;
                            continue;
                        }
                        mine[e * feature_dim + j] += g * w[k];
                    }
                }
            }
        }
        float* out = grad_codebook + lvl_begin * feature_dim;
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < span; ++e) {
            float acc = 0.0f;
            for (int t = 0; t < nthreads; ++t) acc += priv[(size_t)span * t + e];
            out[e] = acc;
        }
        free(priv);
    }
    return redirected;
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the timed reference arm asks for all host cores explicitly. */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
