// tiled_kernels.cuh -- the B200 fast path: spatially tiled fused latent grid kernels.
//
// Why: measured on B200 (profiles/r01_microbench_b200.csv) random 4-byte gathers from an L2-resident
// table run at ~300 Glanes/s and float REDG at ~200 Glanes/s -- collapsing to ~4 Glanes/s when a
// few hundred addresses take all the traffic (coarse levels) -- while shared memory serves random
// gathers at ~2300 Glanes/s and *integer* shared atomics at ~1600 Glanes/s irrespective of conflicts.
// So points are binned into spatial tiles once per coordinate set (the "plan"), and one CTA per
// tile stages the grid NODES its tile touches (about 2-3 per point, all levels together, instead
// of 2^D * L gathers per point) in shared memory:
//   forward : nodes <- rint(latents) once per tile; every point lerps from shared memory, applies
//             the affine decode and writes its output row at its ORIGINAL index;
//   backward: every point adds w_k * (A^T g) to its tile's node accumulators in shared memory as
//             fixed-point integers (exact, order-independent sums; per-tile per-level power-of-two
//             scale from the tile's max |gradient|), then the tile flushes each touched node with ONE
//             float REDG. Decoder gradients use warp REDUX on the same fixed-point values.
// Levels whose node box does not fit the shared-memory budget (fine 3D levels, huge resolutions)
// fall back per level to direct global gathers / REDG inside the same kernel.
//
// Indices, weights and rounding are computed by the same device functions as the point-parallel
// kernels (common.cuh), so hash indices and quantised latents stay bit-identical to the reference.
#pragma once
#include "common.cuh"

namespace shacira {

constexpr int kTileThreads = 128;
constexpr int kMaxTiles = 4096;
constexpr int kBatch = 2048;  // points accumulated between two flushes of a tile (bounds the fixed-point sums)

struct PlanView {
    const int32_t* perm;          // sorted position -> original point index
    const float* coords_sorted;   // [n, D] coordinates in sorted order
    const int32_t* tile_off;      // [ntiles + 1]
    int64_t n;
    int32_t g;                    // tiles per axis (power of two)
    int32_t ntiles;
};

// ---------------------------------------------------------------------------------------------
// plan construction: tile id -> histogram -> scan -> scatter
// ---------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ int tile_of(const float* __restrict__ coords, int64_t i, int g) {
    int id = 0, mul = 1;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const double t = unit_coord(__ldg(coords + i * D + d));
        int ti = (int)floor(t * (double)g);
        ti = max(0, min(g - 1, ti));
        id += ti * mul;
        mul *= g;
    }
    return id;
}

template <int D>
__global__ void __launch_bounds__(1024)
plan_count_kernel(const float* __restrict__ coords, int64_t n, int g, int ntiles, int32_t* __restrict__ tile_id,
                  int32_t* __restrict__ counts) {
    extern __shared__ int s_hist[];
    for (int e = threadIdx.x; e < ntiles; e += blockDim.x) s_hist[e] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = tile_of<D>(coords, i, g);
        tile_id[i] = t;
        atomicAdd(&s_hist[t], 1);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ntiles; e += blockDim.x)
        if (s_hist[e]) atomicAdd(&counts[e], s_hist[e]);
}

// single block: exclusive scan of counts[ntiles] -> tile_off[ntiles+1]; cursor = tile_off
__global__ void __launch_bounds__(1024)
plan_scan_kernel(const int32_t* __restrict__ counts, int ntiles, int32_t* __restrict__ tile_off,
                 int32_t* __restrict__ cursor) {
    __shared__ int s_part[1024];
    const int per = (ntiles + 1023) / 1024;
    const int b = threadIdx.x * per;
    int sum = 0;
    for (int k = 0; k < per; ++k)
        if (b + k < ntiles) sum += counts[b + k];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
        const int v = (threadIdx.x >= o) ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = s_part[threadIdx.x] - sum;
    for (int k = 0; k < per; ++k)
        if (b + k < ntiles) {
            tile_off[b + k] = run;
            cursor[b + k] = run;
            run += counts[b + k];
        }
    if (threadIdx.x == 1023) tile_off[ntiles] = s_part[1023];
}

template <int D>
__global__ void __launch_bounds__(1024)
plan_scatter_kernel(const float* __restrict__ coords, int64_t n, int ntiles, const int32_t* __restrict__ tile_id,
                    int32_t* __restrict__ cursor, int32_t* __restrict__ perm, float* __restrict__ coords_sorted) {
    extern __shared__ int s_mem[];
    int* s_hist = s_mem;            // local count per tile, then global base of this block's run
    for (int e = threadIdx.x; e < ntiles; e += blockDim.x) s_hist[e] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int t = 0, rank = 0;
    if (i < n) {
        t = tile_id[i];
        rank = atomicAdd(&s_hist[t], 1);  // rank of this point among the block's points of tile t
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ntiles; e += blockDim.x) {
        const int c = s_hist[e];
        if (c) s_hist[e] = atomicAdd(&cursor[e], c);  // reserve a run in the tile's segment
    }
    __syncthreads();
    if (i < n) {
        const int pos = s_hist[t] + rank;
        perm[pos] = (int32_t)i;
#pragma unroll
        for (int d = 0; d < D; ++d) coords_sorted[(int64_t)pos * D + d] = __ldg(coords + i * D + d);
    }
}

// ---------------------------------------------------------------------------------------------
// per-tile level geometry
// ---------------------------------------------------------------------------------------------
template <int D>
struct TileGeom {
    int c0[SHACIRA_MAX_LEVELS][D];  // first cell of the tile's box per level/axis
    int w[SHACIRA_MAX_LEVELS][D];   // nodes per axis (cells + 1)
    int off[SHACIRA_MAX_LEVELS + 1];
    unsigned staged;                // bit l: level l lives in shared memory for this tile
    int total;
};

// Cells reachable by points of tile axis-interval [ti/g, (ti+1)/g]: locate() is monotone in t.
template <int D>
__device__ __forceinline__ void tile_geometry(TileGeom<D>& tg, const LevelParams& lp, const int (&ti)[D], int g,
                                              int cap) {
    const int l = threadIdx.x;
    if (l < lp.num_lods) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            int ca, cb;
            float f, gg;
            locate((double)ti[d] / (double)g, lp.res[l], lp.hi[l], ca, f, gg);
            locate((double)(ti[d] + 1) / (double)g, lp.res[l], lp.hi[l], cb, f, gg);
            tg.c0[l][d] = ca;
            tg.w[l][d] = cb - ca + 2;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        unsigned staged = 0;
        for (int q = 0; q < lp.num_lods; ++q) {
            long long nodes = 1;
#pragma unroll
            for (int d = 0; d < D; ++d) nodes *= tg.w[q][d];
            tg.off[q] = run;
            if (run + nodes <= cap) {
                staged |= 1u << q;
                run += (int)nodes;
            }
        }
        tg.off[lp.num_lods] = run;
        tg.staged = staged;
        tg.total = run;
    }
    __syncthreads();
}

// node e of staged level l -> (level-local table row, or -1 when the node lies outside the level)
template <int D>
__device__ __forceinline__ int node_row(const TileGeom<D>& tg, const LevelParams& lp, int l, int e, bool clamp_inside) {
    const int res = lp.res[l];
    int nx[D];
    int r = e;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        nx[d] = tg.c0[l][d] + r % tg.w[l][d];
        r /= tg.w[l][d];
    }
    if ((lp.dense_mask >> l) & 1u) {
        long long idx = nx[0];
        long long mul = res;
        bool outside = nx[0] > res;  // nodes past the grid only ever carry weight 0
#pragma unroll
        for (int d = 1; d < D; ++d) {
            idx += (long long)nx[d] * mul;
            mul *= res;
            outside |= nx[d] > res;
        }
        if (idx >= lp.rows[l] || outside) return clamp_inside ? lp.rows[l] - 1 : -1;
        return (int)idx;
    }
    uint32_t h = (uint32_t)nx[0];
    if (D > 1) h ^= (uint32_t)nx[1] * kPrimeY;
    if (D > 2) h ^= (uint32_t)nx[2] * kPrimeZ;
    return (int)(h & lp.hash_mask);
}

template <int D>
__device__ __forceinline__ int level_of_node(const TileGeom<D>& tg, int L, int e) {
    int a = 0, b = L;  // last level with off <= e (unstaged levels have zero extent)
    while (b - a > 1) {
        const int m = (a + b) >> 1;
        if (tg.off[m] <= e) a = m; else b = m;
    }
    // skip back over unstaged (empty) levels that share the same offset
    while (a > 0 && !((tg.staged >> a) & 1u)) --a;
    return a;
}

// cell, local node index and weights of one point at one level
template <int D>
struct LocalCorners {
    int base;          // local index of corner 0 in the tile's node box (staged levels)
    int stride[D];     // local index step per axis
    float w[1 << D];
};

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int D, int C, int F>
__global__ void __launch_bounds__(kTileThreads)
latent_fwd_tiled_kernel(const PlanView pv, const float* __restrict__ latents, const __grid_constant__ LevelParams lp,
                        const float* __restrict__ A, const float* __restrict__ shift, int per_level, int round_flag,
                        float* __restrict__ feats, int cap) {
    extern __shared__ float s_dyn[];
    __shared__ TileGeom<D> tg;
    const int L = lp.num_lods;
    const int nA = per_level ? L : 1;
    float* s_nodes = s_dyn;                  // [cap][C]
    float* s_A = s_dyn + (size_t)cap * C;    // [nA][C][F]
    float* s_shift = s_A + nA * C * F;       // [nA][F]
    const int tile = blockIdx.x;
    const int beg = pv.tile_off[tile], end = pv.tile_off[tile + 1];
    if (beg == end) return;
    int ti[D];
    {
        int r = tile;
#pragma unroll
        for (int d = 0; d < D; ++d) { ti[d] = r % pv.g; r /= pv.g; }
    }
    for (int e = threadIdx.x; e < nA * C * F; e += kTileThreads) s_A[e] = A[e];
    for (int e = threadIdx.x; e < nA * F; e += kTileThreads) s_shift[e] = shift ? shift[e] : 0.0f;
    tile_geometry<D>(tg, lp, ti, pv.g, cap);

    // stage the tile's nodes: one gather per node instead of 2^D per point
    for (int e = threadIdx.x; e < tg.total; e += kTileThreads) {
        const int l = level_of_node<D>(tg, L, e);
        const int row = node_row<D>(tg, lp, l, e - tg.off[l], true);
        float v[C];
        load_row<C>(latents + ((int64_t)lp.first[l] + row) * C, v);
#pragma unroll
        for (int ch = 0; ch < C; ++ch) s_nodes[(size_t)e * C + ch] = round_flag ? rintf(v[ch]) : v[ch];
    }
    __syncthreads();

    constexpr int NC = 1 << D;
    constexpr int G = (F >= 4) ? 1 : 4 / F;  // levels per 16-byte output vector
    const bool vec_o = (L * F) % 4 == 0 && (L % G) == 0;
    for (int j = beg + threadIdx.x; j < end; j += kTileThreads) {
        double t[D];
        load_unit_coords<D>(pv.coords_sorted, j, t);
        float* out = feats + (int64_t)__ldg(pv.perm + j) * L * F;
        float o[G * F];
        for (int l = 0; l < L; ++l) {
            float z[C];
            if ((tg.staged >> l) & 1u) {
                const int32_t res = lp.res[l];
                const float hi = lp.hi[l];
                int p[D];
                float f[D], g1[D];
#pragma unroll
                for (int d = 0; d < D; ++d) locate(t[d], res, hi, p[d], f[d], g1[d]);
                int base = tg.off[l], stride = 1, st[D];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    base += (p[d] - tg.c0[l][d]) * stride;
                    st[d] = stride;
                    stride *= tg.w[l][d];
                }
                float w[NC];
                int li[NC];
                if constexpr (D == 2) {
                    w[0] = __fmul_rn(g1[0], g1[1]); w[1] = __fmul_rn(g1[0], f[1]);
                    w[2] = __fmul_rn(f[0], g1[1]);  w[3] = __fmul_rn(f[0], f[1]);
                    li[0] = base; li[1] = base + st[1]; li[2] = base + st[0]; li[3] = base + st[0] + st[1];
                } else {
                    const float gg = __fmul_rn(g1[0], g1[1]), gf = __fmul_rn(g1[0], f[1]);
                    const float fg = __fmul_rn(f[0], g1[1]), ff = __fmul_rn(f[0], f[1]);
                    w[0] = __fmul_rn(gg, g1[2]); w[1] = __fmul_rn(gg, f[2]); w[2] = __fmul_rn(gf, g1[2]);
                    w[3] = __fmul_rn(gf, f[2]);  w[4] = __fmul_rn(fg, g1[2]); w[5] = __fmul_rn(fg, f[2]);
                    w[6] = __fmul_rn(ff, g1[2]); w[7] = __fmul_rn(ff, f[2]);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        li[k] = base + ((k >> 2) & 1) * st[0] + ((k >> 1) & 1) * st[1] + (k & 1) * st[2];
                }
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    // same contraction order as the reference build: fma(v0,w0, v1*w1), then k = 2..
                    float acc = __fmul_rn(s_nodes[(size_t)li[1] * C + ch], w[1]);
                    acc = __fmaf_rn(s_nodes[(size_t)li[0] * C + ch], w[0], acc);
#pragma unroll
                    for (int k = 2; k < NC; ++k) acc = __fmaf_rn(s_nodes[(size_t)li[k] * C + ch], w[k], acc);
                    z[ch] = acc;
                }
            } else {  // level too large for the tile's shared-memory box: direct global gathers
                Corners<D> c;
                corners<D>(t, lp, l, c);
                const float* base = latents + (int64_t)lp.first[l] * C;
                float v[NC][C];
#pragma unroll
                for (int k = 0; k < NC; ++k) load_row<C>(base + (int64_t)c.idx[k] * C, v[k]);
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    float acc = __fmul_rn(round_flag ? rintf(v[1][ch]) : v[1][ch], c.w[1]);
                    acc = __fmaf_rn(round_flag ? rintf(v[0][ch]) : v[0][ch], c.w[0], acc);
#pragma unroll
                    for (int k = 2; k < NC; ++k) acc = __fmaf_rn(round_flag ? rintf(v[k][ch]) : v[k][ch], c.w[k], acc);
                    z[ch] = acc;
                }
            }
            const int la = per_level ? l : 0;
            const int q = l % G;
#pragma unroll
            for (int jf = 0; jf < F; ++jf) {
                float acc = s_shift[la * F + jf];
#pragma unroll
                for (int ch = 0; ch < C; ++ch) acc = __fmaf_rn(z[ch], s_A[(la * C + ch) * F + jf], acc);
                if (vec_o) {
#pragma unroll
                    for (int qq = 0; qq < G; ++qq)
                        if (qq == q) o[qq * F + jf] = acc;
                } else {
                    out[l * F + jf] = acc;
                }
            }
            if (vec_o && q == G - 1) store_row<G * F>(out + (l - (G - 1)) * F, o);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// Power-of-two scale so that the sum of up to 2^k terms of magnitude <= m stays below 2^30.
__device__ __forceinline__ float fixed_scale(float m, int k, float& inv) {
    if (!(m > 0.0f) || !isfinite(m)) {
        inv = 0.0f;
        return 0.0f;
    }
    int ex;
    frexpf(m, &ex);  // m < 2^ex
    const int e = 30 - k - ex;
    inv = ldexpf(1.0f, -e);
    return ldexpf(1.0f, e);
}

template <int D, int C, int F, bool DEC>
__global__ void __launch_bounds__(kTileThreads)
latent_bwd_tiled_kernel(const PlanView pv, const float* __restrict__ grad_out, const float* __restrict__ latents,
                        const __grid_constant__ LevelParams lp, const float* __restrict__ A, int per_level,
                        int round_flag, float* __restrict__ grad_latents, float* __restrict__ grad_A,
                        float* __restrict__ grad_shift, int cap) {
    extern __shared__ float s_dyn[];
    __shared__ TileGeom<D> tg;
    __shared__ unsigned s_gmax[SHACIRA_MAX_LEVELS];   // max |A^T g| per level over the batch (float bits)
    __shared__ unsigned s_omax[SHACIRA_MAX_LEVELS];   // max |g| per level (decoder-gradient scale)
    __shared__ unsigned s_zmax;                        // max |staged latent|
    __shared__ float s_scale[SHACIRA_MAX_LEVELS], s_inv[SHACIRA_MAX_LEVELS];
    __shared__ float s_oscale[SHACIRA_MAX_LEVELS], s_oinv[SHACIRA_MAX_LEVELS];
    const int L = lp.num_lods;
    const int nA = per_level ? L : 1;
    constexpr int NW = kTileThreads / 32;
    int* s_acc = reinterpret_cast<int*>(s_dyn);                           // [cap][C] fixed-point node sums
    float* s_lat = s_dyn + (size_t)cap * C;                               // [cap][C] staged latents (DEC)
    float* s_A = s_lat + (DEC ? (size_t)cap * C : 0);                     // [nA][C][F]
    float* s_gA = s_A + nA * C * F;                                       // [NW][L][C][F] per-warp partial sums
    float* s_gS = s_gA + (DEC ? NW * L * C * F : 0);                      // [NW][L][F]
    const int tile = blockIdx.x;
    const int beg = pv.tile_off[tile], end = pv.tile_off[tile + 1];
    if (beg == end) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ti[D];
    {
        int r = tile;
#pragma unroll
        for (int d = 0; d < D; ++d) { ti[d] = r % pv.g; r /= pv.g; }
    }
    for (int e = threadIdx.x; e < nA * C * F; e += kTileThreads) s_A[e] = A[e];
    if (DEC)
        for (int e = threadIdx.x; e < NW * L * (C * F + F); e += kTileThreads) s_gA[e] = 0.0f;
    if (threadIdx.x == 0) s_zmax = 0u;
    tile_geometry<D>(tg, lp, ti, pv.g, cap);
    if (DEC) {
        unsigned zm = 0u;
        for (int e = threadIdx.x; e < tg.total; e += kTileThreads) {
            const int l = level_of_node<D>(tg, L, e);
            const int row = node_row<D>(tg, lp, l, e - tg.off[l], true);
            float v[C];
            load_row<C>(latents + ((int64_t)lp.first[l] + row) * C, v);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                const float q = round_flag ? rintf(v[ch]) : v[ch];
                s_lat[(size_t)e * C + ch] = q;
                zm = max(zm, __float_as_uint(fabsf(q)));
            }
        }
        zm = __reduce_max_sync(0xffffffffu, zm);
        if (lane == 0) atomicMax(&s_zmax, zm);
    }
    constexpr int NC = 1 << D;
    const bool vec_g = (F == 1) || ((L * F) % (F >= 4 ? 4 : F) == 0);

    for (int b0 = beg; b0 < end; b0 += kBatch) {
        const int b1 = min(end, b0 + kBatch);
        int kbits = 0;
        while ((1 << kbits) < (b1 - b0)) ++kbits;
        for (int e = threadIdx.x; e < tg.total * C; e += kTileThreads) s_acc[e] = 0;
        if (threadIdx.x < SHACIRA_MAX_LEVELS) { s_gmax[threadIdx.x] = 0u; s_omax[threadIdx.x] = 0u; }
        __syncthreads();
        // pass 1: per-level maxima of what will be accumulated
        for (int j0 = b0; j0 < b1; j0 += kTileThreads) {
            const int j = j0 + threadIdx.x;
            const bool live = j < b1;
            const float* g_row = grad_out + (int64_t)(live ? __ldg(pv.perm + j) : 0) * L * F;
            for (int l = 0; l < L; ++l) {
                float g[F];
#pragma unroll
                for (int jf = 0; jf < F; ++jf) g[jf] = 0.0f;
                if (live) {
                    if (vec_g) load_row<F>(g_row + l * F, g);
                    else {
#pragma unroll
                        for (int jf = 0; jf < F; ++jf) g[jf] = __ldg(g_row + l * F + jf);
                    }
                }
                const int la = per_level ? l : 0;
                float m = 0.0f, mo = 0.0f;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    float acc = 0.0f;
#pragma unroll
                    for (int jf = 0; jf < F; ++jf) acc = __fmaf_rn(g[jf], s_A[(la * C + ch) * F + jf], acc);
                    m = fmaxf(m, fabsf(acc));
                }
#pragma unroll
                for (int jf = 0; jf < F; ++jf) mo = fmaxf(mo, fabsf(g[jf]));
                const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
                if (lane == 0 && wm) atomicMax(&s_gmax[l], wm);
                if (DEC) {
                    const unsigned wo = __reduce_max_sync(0xffffffffu, __float_as_uint(mo));
                    if (lane == 0 && wo) atomicMax(&s_omax[l], wo);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < L) {
            float inv;
            s_scale[threadIdx.x] = fixed_scale(__uint_as_float(s_gmax[threadIdx.x]), kbits, inv);
            s_inv[threadIdx.x] = inv;
            if (DEC) {
                s_oscale[threadIdx.x] = fixed_scale(__uint_as_float(s_omax[threadIdx.x]), 5, inv);  // 32-lane REDUX
                s_oinv[threadIdx.x] = inv;
            }
        }
        __syncthreads();
        float zscale = 0.0f, zinv = 0.0f;
        if (DEC) {
            // z is a convex combination of staged values: |z| <= zmax (direct levels use their own bound below)
            int ex = 0;
            const float zm = __uint_as_float(s_zmax);
            if (zm > 0.0f) frexpf(zm, &ex);
            zscale = ldexpf(1.0f, -ex);  // z * zscale in [-1, 1]
            zinv = ldexpf(1.0f, ex);
        }
        // pass 2: accumulate
        for (int j0 = b0; j0 < b1; j0 += kTileThreads) {
            const int j = j0 + threadIdx.x;
            const bool live = j < b1;
            double t[D];
            if (live) load_unit_coords<D>(pv.coords_sorted, j, t);
            const float* g_row = grad_out + (int64_t)(live ? __ldg(pv.perm + j) : 0) * L * F;
            for (int l = 0; l < L; ++l) {
                float g[F];
#pragma unroll
                for (int jf = 0; jf < F; ++jf) g[jf] = 0.0f;
                float z[C];
#pragma unroll
                for (int ch = 0; ch < C; ++ch) z[ch] = 0.0f;
                const int la = per_level ? l : 0;
                if (live) {
                    if (vec_g) load_row<F>(g_row + l * F, g);
                    else {
#pragma unroll
                        for (int jf = 0; jf < F; ++jf) g[jf] = __ldg(g_row + l * F + jf);
                    }
                    float gz[C];
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) {
                        float acc = 0.0f;
#pragma unroll
                        for (int jf = 0; jf < F; ++jf) acc = __fmaf_rn(g[jf], s_A[(la * C + ch) * F + jf], acc);
                        gz[ch] = acc;
                    }
                    if ((tg.staged >> l) & 1u) {
                        const int32_t res = lp.res[l];
                        const float hi = lp.hi[l];
                        int p[D];
                        float f[D], g1[D];
#pragma unroll
                        for (int d = 0; d < D; ++d) locate(t[d], res, hi, p[d], f[d], g1[d]);
                        int base = tg.off[l], stride = 1, st[D];
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            base += (p[d] - tg.c0[l][d]) * stride;
                            st[d] = stride;
                            stride *= tg.w[l][d];
                        }
                        float w[NC];
                        int li[NC];
                        if constexpr (D == 2) {
                            w[0] = __fmul_rn(g1[0], g1[1]); w[1] = __fmul_rn(g1[0], f[1]);
                            w[2] = __fmul_rn(f[0], g1[1]);  w[3] = __fmul_rn(f[0], f[1]);
                            li[0] = base; li[1] = base + st[1]; li[2] = base + st[0]; li[3] = base + st[0] + st[1];
                        } else {
                            const float gg = __fmul_rn(g1[0], g1[1]), gf = __fmul_rn(g1[0], f[1]);
                            const float fg = __fmul_rn(f[0], g1[1]), ff = __fmul_rn(f[0], f[1]);
                            w[0] = __fmul_rn(gg, g1[2]); w[1] = __fmul_rn(gg, f[2]); w[2] = __fmul_rn(gf, g1[2]);
                            w[3] = __fmul_rn(gf, f[2]);  w[4] = __fmul_rn(fg, g1[2]); w[5] = __fmul_rn(fg, f[2]);
                            w[6] = __fmul_rn(ff, g1[2]); w[7] = __fmul_rn(ff, f[2]);
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                li[k] = base + ((k >> 2) & 1) * st[0] + ((k >> 1) & 1) * st[1] + (k & 1) * st[2];
                        }
                        const float sc = s_scale[l];
#pragma unroll
                        for (int k = 0; k < NC; ++k) {
#pragma unroll
                            for (int ch = 0; ch < C; ++ch) {
                                const int q = __float2int_rn(__fmul_rn(__fmul_rn(gz[ch], w[k]), sc));
                                if (q) atomicAdd(&s_acc[(size_t)li[k] * C + ch], q);
                            }
                        }
                        if (DEC) {
#pragma unroll
                            for (int ch = 0; ch < C; ++ch) {
                                float acc = __fmul_rn(s_lat[(size_t)li[1] * C + ch], w[1]);
                                acc = __fmaf_rn(s_lat[(size_t)li[0] * C + ch], w[0], acc);
#pragma unroll
                                for (int k = 2; k < NC; ++k) acc = __fmaf_rn(s_lat[(size_t)li[k] * C + ch], w[k], acc);
                                z[ch] = acc;
                            }
                        }
                    } else {  // direct level: float REDG to global, gathers for z
                        Corners<D> c;
                        corners<D>(t, lp, l, c);
                        float* base = grad_latents + (int64_t)lp.first[l] * C;
#pragma unroll
                        for (int k = 0; k < NC; ++k) {
                            float gv[C];
#pragma unroll
                            for (int ch = 0; ch < C; ++ch) gv[ch] = __fmul_rn(gz[ch], c.w[k]);
                            red_add_row<C>(base + (int64_t)c.idx[k] * C, gv);
                        }
                        if (DEC) {
                            const float* lb = latents + (int64_t)lp.first[l] * C;
#pragma unroll
                            for (int k = 0; k < NC; ++k) {
                                float v[C];
                                load_row<C>(lb + (int64_t)c.idx[k] * C, v);
#pragma unroll
                                for (int ch = 0; ch < C; ++ch)
                                    z[ch] = __fmaf_rn(round_flag ? rintf(v[ch]) : v[ch], c.w[k], z[ch]);
                            }
                        }
                    }
                }
                if (DEC) {
                    // decoder gradients: fixed-point REDUX over the warp, one float add per warp and value
                    const float os = s_oscale[l], oi = s_oinv[l];
                    const bool staged_l = (tg.staged >> l) & 1u;
#pragma unroll
                    for (int jf = 0; jf < F; ++jf) {
                        const int qs = __float2int_rn(__fmul_rn(g[jf], os));
                        const int ss = __reduce_add_sync(0xffffffffu, qs);
                        if (lane == 0 && ss) s_gS[(warp * L + l) * F + jf] += (float)ss * oi;
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) {
                            if (staged_l) {
                                const int qa = __float2int_rn(__fmul_rn(__fmul_rn(z[ch], zscale), __fmul_rn(g[jf], os)));
                                const int sa = __reduce_add_sync(0xffffffffu, qa);
                                if (lane == 0 && sa) s_gA[((warp * L + l) * C + ch) * F + jf] += (float)sa * oi * zinv;
                            } else {
                                const float sa = warp_sum(z[ch] * g[jf]);
                                if (lane == 0) s_gA[((warp * L + l) * C + ch) * F + jf] += sa;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        // flush: one float REDG per touched node
        for (int e = threadIdx.x; e < tg.total; e += kTileThreads) {
            const int l = level_of_node<D>(tg, L, e);
            bool any = false;
            float gv[C];
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                const int q = s_acc[(size_t)e * C + ch];
                any |= (q != 0);
                gv[ch] = (float)q * s_inv[l];
            }
            if (any) {
                const int row = node_row<D>(tg, lp, l, e - tg.off[l], false);
                if (row >= 0) red_add_row<C>(grad_latents + ((int64_t)lp.first[l] + row) * C, gv);
            }
        }
        __syncthreads();
    }
    if (DEC) {
        for (int e = threadIdx.x; e < L * C * F; e += kTileThreads) {
            float s = 0.0f;
#pragma unroll
            for (int wq = 0; wq < NW; ++wq) s += s_gA[wq * L * C * F + e];
            if (grad_A && s != 0.0f) red_add(grad_A + e, s);
        }
        for (int e = threadIdx.x; e < L * F; e += kTileThreads) {
            float s = 0.0f;
#pragma unroll
            for (int wq = 0; wq < NW; ++wq) s += s_gS[wq * L * F + e];
            if (grad_shift && s != 0.0f) red_add(grad_shift + e, s);
        }
    }
}

}  // namespace shacira
