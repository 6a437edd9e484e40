// common.cuh -- level metadata, coordinate math and spatial hash shared by every kernel.
//
// The arithmetic restates wisp/csrc/ops/hashgrid_interpolate2d_cuda.cu:17-36,62-88 and
// wisp/csrc/ops/hashgrid_interpolate_cuda.cu:17-39,66-95 of the reference so that cell
// positions, corner indices and weights are bit-identical (see DESIGN.md "Arithmetic").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/shacira_b200.h"

namespace shacira {

constexpr uint32_t kPrimeY = 2654435761u;  // hashgrid_interpolate2d_cuda.cu:25
constexpr uint32_t kPrimeZ = 805459861u;   // hashgrid_interpolate_cuda.cu:25

// Passed by value as a __grid_constant__ kernel parameter (lives in constant bank 0).
struct LevelParams {
    double resd[SHACIRA_MAX_LEVELS];    // (double)res: x = float(resd * t) without a per-level conversion
    int32_t res[SHACIRA_MAX_LEVELS];    // grid resolution of the level
    int32_t first[SHACIRA_MAX_LEVELS];  // first table row of the level
    int32_t rows[SHACIRA_MAX_LEVELS];   // rows of the level = min(2^bw, res^dim)
    float hi[SHACIRA_MAX_LEVELS];       // (float)(res - 1 - 1e-5): upper clamp bound
    uint32_t dense_mask;                // bit l set: level l uses x + y*res (+ z*res^2)
    uint32_t hash_mask;                 // 2^bw - 1
    int32_t num_lods;
    int32_t pad;
};

// t = coord * 0.5 + 0.5 in double. The reference evaluates resolution * (coord*0.5+0.5)
// in double (2d_cuda.cu:65); the first two operations do not depend on the level.
__device__ __forceinline__ double unit_coord(float c) { return fma((double)c, 0.5, 0.5); }

// x = clamp((float)(res * t), 0, hi); cell = floor(x); f = x - cell; g = 1 - f.
// (1.0 - f is a double subtraction narrowed to float in the reference; it is exact in
// double for every reachable f, hence equal to the float subtraction -- DESIGN.md.)
// floor(x) and float(floor(x)) for 0 <= x < 2^22 without the conversion unit: x + 2^23 rounded DOWN is exactly
// 2^23 + floor(x) (floats in [2^23, 2^24) are the integers), whose mantissa is the cell and which minus 2^23 is the
// cell as a float. F2I / I2F run on the quarter-rate XU pipe with scoreboard latency; FADD.RM is a plain FMA-pipe op.
// (Resolutions are capped at 2^22 by build_levels.)
__device__ __forceinline__ void floor_cell(float x, int32_t& cell, float& cellf) {
    const float y = __fadd_rd(x, 8388608.0f);
    cell = __float_as_int(y) - 0x4B000000;
    cellf = __fsub_rn(y, 8388608.0f);
}
__device__ __forceinline__ void locate(double t, int32_t res, float hi, int32_t& cell, float& f, float& g) {
    float x = __double2float_rn(__dmul_rn((double)res, t));
    x = fmaxf(0.0f, fminf(hi, x));
    float cf;
    floor_cell(x, cell, cf);
    f = __fsub_rn(x, cf);
    g = __fsub_rn(1.0f, f);
}

template <int D>
struct Corners {
    int32_t idx[1 << D];  // level-local row index
    float w[1 << D];
};

// Corner j of the reference: 2D x += (j>>1)&1, y += j&1 (2d_cuda.cu:83-88);
// 3D x += (j>>2)&1, y += (j>>1)&1, z += j&1 (_cuda.cu:88-94). Weights 2d_cuda.cu:72-75,
// _cuda.cu:77-84 (products evaluated left to right).
template <int D>
__device__ __forceinline__ void corners(const double (&t)[D], const LevelParams& lp, int l, Corners<D>& c) {
    const int32_t res = lp.res[l];
    const float hi = lp.hi[l];
    const bool dense = (lp.dense_mask >> l) & 1u;
    int32_t p[D];
    float f[D], g[D];
#pragma unroll
    for (int d = 0; d < D; ++d) locate(t[d], res, hi, p[d], f[d], g[d]);
    if constexpr (D == 2) {
        c.w[0] = __fmul_rn(g[0], g[1]);
        c.w[1] = __fmul_rn(g[0], f[1]);
        c.w[2] = __fmul_rn(f[0], g[1]);
        c.w[3] = __fmul_rn(f[0], f[1]);
        if (dense) {
            // Reference Q4: on dense levels with res >= 257 the clamp bound equals res-1, so a
            // corner can be res (one past the level) with weight exactly 0. Keep the row index
            // inside the level; the product with the zero weight is unchanged.
            const int32_t last = lp.rows[l] - 1;
            const int32_t base = p[0] + p[1] * res;
            c.idx[0] = min(base, last);
            c.idx[1] = min(base + res, last);
            c.idx[2] = min(base + 1, last);
            c.idx[3] = min(base + res + 1, last);
        } else {
            const uint32_t m = lp.hash_mask;
            const uint32_t hx0 = (uint32_t)p[0], hx1 = (uint32_t)p[0] + 1u;
            const uint32_t hy0 = (uint32_t)p[1] * kPrimeY, hy1 = hy0 + kPrimeY;
            c.idx[0] = (int32_t)((hx0 ^ hy0) & m);
            c.idx[1] = (int32_t)((hx0 ^ hy1) & m);
            c.idx[2] = (int32_t)((hx1 ^ hy0) & m);
            c.idx[3] = (int32_t)((hx1 ^ hy1) & m);
        }
    } else {
        const float gg = __fmul_rn(g[0], g[1]), gf = __fmul_rn(g[0], f[1]);
        const float fg = __fmul_rn(f[0], g[1]), ff = __fmul_rn(f[0], f[1]);
        c.w[0] = __fmul_rn(gg, g[2]);
        c.w[1] = __fmul_rn(gg, f[2]);
        c.w[2] = __fmul_rn(gf, g[2]);
        c.w[3] = __fmul_rn(gf, f[2]);
        c.w[4] = __fmul_rn(fg, g[2]);
        c.w[5] = __fmul_rn(fg, f[2]);
        c.w[6] = __fmul_rn(ff, g[2]);
        c.w[7] = __fmul_rn(ff, f[2]);
        if (dense) {
            const int32_t last = lp.rows[l] - 1;
            const int32_t rr = res * res;
            const int32_t base = p[0] + p[1] * res + p[2] * rr;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                c.idx[j] = min(base + ((j >> 2) & 1) + ((j >> 1) & 1) * res + (j & 1) * rr, last);
        } else {
            const uint32_t m = lp.hash_mask;
            const uint32_t hx[2] = {(uint32_t)p[0], (uint32_t)p[0] + 1u};
            const uint32_t hy0 = (uint32_t)p[1] * kPrimeY, hz0 = (uint32_t)p[2] * kPrimeZ;
            const uint32_t hy[2] = {hy0, hy0 + kPrimeY};
            const uint32_t hz[2] = {hz0, hz0 + kPrimeZ};
#pragma unroll
            for (int j = 0; j < 8; ++j) c.idx[j] = (int32_t)((hx[(j >> 2) & 1] ^ hy[(j >> 1) & 1] ^ hz[j & 1]) & m);
        }
    }
}

template <int D>
__device__ __forceinline__ void load_unit_coords(const float* __restrict__ coords, int64_t i, double (&t)[D]) {
    if constexpr (D == 2) {
        const float2 c = __ldg(reinterpret_cast<const float2*>(coords) + i);
        t[0] = unit_coord(c.x);
        t[1] = unit_coord(c.y);
    } else {
        t[0] = unit_coord(__ldg(coords + i * 3 + 0));
        t[1] = unit_coord(__ldg(coords + i * 3 + 1));
        t[2] = unit_coord(__ldg(coords + i * 3 + 2));
    }
}

// N floats from global memory as the widest aligned vector (rows are N*4-byte aligned).
template <int N>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&v)[N]) {
    if constexpr (N == 1) {
        v[0] = __ldg(p);
    } else if constexpr (N == 2) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = a.x;
        v[1] = a.y;
    } else {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p) + q);
            v[4 * q + 0] = a.x;
            v[4 * q + 1] = a.y;
            v[4 * q + 2] = a.z;
            v[4 * q + 3] = a.w;
        }
    }
}

template <int N>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&v)[N]) {
    if constexpr (N == 1) {
        *p = v[0];
    } else if constexpr (N == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int q = 0; q < N / 4; ++q)
            reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}

// Fire-and-forget float adds (REDG.E.ADD.F32); vector forms need sm_90+.
__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <int N>
__device__ __forceinline__ void red_add_row(float* p, const float (&v)[N]) {
    if constexpr (N == 1) {
        red_add(p, v[0]);
    } else if constexpr (N == 2) {
        red_add2(p, v[0], v[1]);
    } else {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) red_add4(p + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}

// max that PROPAGATES NaN (fmaxf returns the non-NaN operand): a NaN upstream gradient has to reach the fixed-point
// scale logic, where it poisons the level like the reference's float atomics would. As an unsigned bit pattern a
// positive NaN compares above +Inf, so REDUX / atomicMax on the bits keep it.
__device__ __forceinline__ float nan_max(float m, float a) { return (a != a) ? a : ((m != m) ? m : fmaxf(m, a)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace shacira
