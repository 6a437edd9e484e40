// hashgrid_kernels.cuh -- point-parallel kernels: all levels of the grid in ONE launch.
//
//   hashgrid_fwd_kernel<D,F>     plain table, reference semantics
//                                (hashgrid_interpolate{,2d}_cuda.cu fwd + the host level loop)
//   hashgrid_bwd_kernel<D,F>     scatter-add of grad_output * weight (bwd kernels of the same files)
//   latent_fwd_kernel<D,C,F>     rint(latent) gather -> lerp -> affine decode, fused
//   latent_bwd_kernel<D,C,F>     decode^T -> scatter-add to latents, + grad of the decoder
//   corners_kernel<D>            indices/weights dump for the bit-exactness tests
//
// One thread owns one point: coordinates are read once (the reference re-reads them per
// level), the level loop is unrolled so 2^D * kUnroll independent gathers are in flight,
// and the thread's output row is written as 16-byte vectors.
#pragma once
#include "common.cuh"

namespace shacira {

constexpr int kBlock = 256;

// ------------------------------------------------------------------------------------
// plain forward
// ------------------------------------------------------------------------------------
template <int D, int F>
__global__ void __launch_bounds__(kBlock)
hashgrid_fwd_kernel(const float* __restrict__ coords, int64_t n, const float* __restrict__ table,
                    const __grid_constant__ LevelParams lp, float* __restrict__ feats) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    double t[D];
    load_unit_coords<D>(coords, i, t);
    const int L = lp.num_lods;
    float* out = feats + i * (int64_t)L * F;
    constexpr int NC = 1 << D;
    constexpr int G = (F >= 4) ? 1 : 4 / F;  // levels per 16-byte output vector
    int l = 0;
    for (; l + G <= L; l += G) {
        float o[G * F];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            Corners<D> c;
            corners<D>(t, lp, l + q, c);
            const float* base = table + (int64_t)lp.first[l + q] * F;
            float v[NC][F];
#pragma unroll
            for (int k = 0; k < NC; ++k) load_row<F>(base + (int64_t)c.idx[k] * F, v[k]);
#pragma unroll
            for (int j = 0; j < F; ++j) {
                // contraction order of the reference build: fma(v0,c0, v1*c1), then k = 2..
                float acc = __fmul_rn(v[1][j], c.w[1]);
                acc = __fmaf_rn(v[0][j], c.w[0], acc);
#pragma unroll
                for (int k = 2; k < NC; ++k) acc = __fmaf_rn(v[k][j], c.w[k], acc);
                o[q * F + j] = acc;
            }
        }
        if ((L * F) % 4 == 0) {
            store_row<G * F>(out + l * F, o);
        } else {
#pragma unroll
            for (int j = 0; j < G * F; ++j) out[l * F + j] = o[j];
        }
    }
    for (; l < L; ++l) {  // tail levels (L not a multiple of G)
        Corners<D> c;
        corners<D>(t, lp, l, c);
        const float* base = table + (int64_t)lp.first[l] * F;
        for (int j = 0; j < F; ++j) {
            float acc = __fmul_rn(__ldg(base + (int64_t)c.idx[1] * F + j), c.w[1]);
            acc = __fmaf_rn(__ldg(base + (int64_t)c.idx[0] * F + j), c.w[0], acc);
#pragma unroll
            for (int k = 2; k < NC; ++k) acc = __fmaf_rn(__ldg(base + (int64_t)c.idx[k] * F + j), c.w[k], acc);
            out[l * F + j] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------
// plain backward
// ------------------------------------------------------------------------------------
template <int D, int F>
__global__ void __launch_bounds__(kBlock)
hashgrid_bwd_kernel(const float* __restrict__ coords, int64_t n, const float* __restrict__ grad_out,
                    const __grid_constant__ LevelParams lp, uint32_t skip_mask, float* __restrict__ grad_table) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    double t[D];
    load_unit_coords<D>(coords, i, t);
    const int L = lp.num_lods;
    const float* g_row = grad_out + i * (int64_t)L * F;
    constexpr int NC = 1 << D;
    const bool vec_ok = (F == 1) || ((L * F) % (F >= 4 ? 4 : F) == 0);
#pragma unroll 2
    for (int l = 0; l < L; ++l) {
        if ((skip_mask >> l) & 1u) continue;  // level accumulated in shared memory by coarse_bwd_kernel
        Corners<D> c;
        corners<D>(t, lp, l, c);
        float g[F];
        if (vec_ok) {
            load_row<F>(g_row + l * F, g);
        } else {
#pragma unroll
            for (int j = 0; j < F; ++j) g[j] = __ldg(g_row + l * F + j);
        }
        float* base = grad_table + (int64_t)lp.first[l] * F;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            float gv[F];
#pragma unroll
            for (int j = 0; j < F; ++j) gv[j] = __fmul_rn(g[j], c.w[k]);  // 2d_cuda.cu:203-205
            red_add_row<F>(base + (int64_t)c.idx[k] * F, gv);
        }
    }
}

// ------------------------------------------------------------------------------------
// fused latent forward: q = rint(latent) -> lerp (C channels) -> A^T z + shift (F channels)
// ------------------------------------------------------------------------------------
template <int D, int C, int F>
__global__ void __launch_bounds__(kBlock)
latent_fwd_kernel(const float* __restrict__ coords, int64_t n, const float* __restrict__ latents,
                  const __grid_constant__ LevelParams lp, const float* __restrict__ A,
                  const float* __restrict__ shift, int per_level, int round_flag, float* __restrict__ feats,
                  float* __restrict__ zsave) {
    extern __shared__ float s_dec[];  // [nA][C*F] then [nA][F]
    const int L = lp.num_lods;
    const int nA = per_level ? L : 1;
    float* s_A = s_dec;
    float* s_shift = s_dec + nA * C * F;
    for (int e = threadIdx.x; e < nA * C * F; e += kBlock) s_A[e] = A[e];
    for (int e = threadIdx.x; e < nA * F; e += kBlock) s_shift[e] = shift ? shift[e] : 0.0f;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    double t[D];
    load_unit_coords<D>(coords, i, t);
    float* out = feats + i * (int64_t)L * F;
    float* zout = zsave ? zsave + i * (int64_t)L * C : nullptr;
    constexpr int NC = 1 << D;
    constexpr int G = (F >= 4) ? 1 : 4 / F;
    constexpr int GZ = (C >= 4) ? 1 : 4 / C;
    static_assert(G % GZ == 0 || GZ % G == 0, "level grouping");
    constexpr int GG = (G > GZ) ? G : GZ;  // levels handled per outer iteration
    int l = 0;
    const bool vec_o = (L * F) % 4 == 0, vec_z = (L * C) % 4 == 0;
    for (; l + GG <= L; l += GG) {
        float o[GG * F];
        float zz[GG * C];
#pragma unroll
        for (int q = 0; q < GG; ++q) {
            Corners<D> c;
            corners<D>(t, lp, l + q, c);
            const float* base = latents + (int64_t)lp.first[l + q] * C;
            float v[NC][C];
#pragma unroll
            for (int k = 0; k < NC; ++k) load_row<C>(base + (int64_t)c.idx[k] * C, v[k]);
            const int la = per_level ? (l + q) : 0;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                // torch.round == half-to-even == rintf; contraction order of the reference build
                float acc = __fmul_rn(round_flag ? rintf(v[1][ch]) : v[1][ch], c.w[1]);
                acc = __fmaf_rn(round_flag ? rintf(v[0][ch]) : v[0][ch], c.w[0], acc);
#pragma unroll
                for (int k = 2; k < NC; ++k) acc = __fmaf_rn(round_flag ? rintf(v[k][ch]) : v[k][ch], c.w[k], acc);
                zz[q * C + ch] = acc;
            }
#pragma unroll
            for (int j = 0; j < F; ++j) {
                float acc = s_shift[la * F + j];
#pragma unroll
                for (int ch = 0; ch < C; ++ch) acc = __fmaf_rn(zz[q * C + ch], s_A[(la * C + ch) * F + j], acc);
                o[q * F + j] = acc;
            }
        }
        if (vec_o) {
#pragma unroll
            for (int q = 0; q < GG / G; ++q) {
                float tmp[G * F];
#pragma unroll
                for (int j = 0; j < G * F; ++j) tmp[j] = o[q * G * F + j];
                store_row<G * F>(out + (l + q * G) * F, tmp);
            }
        } else {
#pragma unroll
            for (int j = 0; j < GG * F; ++j) out[l * F + j] = o[j];
        }
        if (zout) {
            if (vec_z) {
#pragma unroll
                for (int q = 0; q < GG / GZ; ++q) {
                    float tmp[GZ * C];
#pragma unroll
                    for (int j = 0; j < GZ * C; ++j) tmp[j] = zz[q * GZ * C + j];
                    store_row<GZ * C>(zout + (l + q * GZ) * C, tmp);
                }
            } else {
#pragma unroll
                for (int j = 0; j < GG * C; ++j) zout[l * C + j] = zz[j];
            }
        }
    }
    for (; l < L; ++l) {  // tail levels
        Corners<D> c;
        corners<D>(t, lp, l, c);
        const float* base = latents + (int64_t)lp.first[l] * C;
        const int la = per_level ? l : 0;
        float zc[C];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            float raw[NC];
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                raw[k] = __ldg(base + (int64_t)c.idx[k] * C + ch);
                if (round_flag) raw[k] = rintf(raw[k]);
            }
            float acc = __fmul_rn(raw[1], c.w[1]);
            acc = __fmaf_rn(raw[0], c.w[0], acc);
#pragma unroll
            for (int k = 2; k < NC; ++k) acc = __fmaf_rn(raw[k], c.w[k], acc);
            zc[ch] = acc;
            if (zout) zout[l * C + ch] = acc;
        }
        for (int j = 0; j < F; ++j) {
            float acc = s_shift[la * F + j];
#pragma unroll
            for (int ch = 0; ch < C; ++ch) acc = __fmaf_rn(zc[ch], s_A[(la * C + ch) * F + j], acc);
            out[l * F + j] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------
// fused latent backward
//   gz[c]            = sum_f g[f] * A[la,c,f]
//   grad_latents[..] += w_k * gz[c]                (straight-through rounding)
//   grad_A[l,c,f]    += z[i,l,c] * g[i,l,f]         grad_shift[l,f] += g[i,l,f]
// Decoder gradients are reduced warp -> block (shared) -> global, one add per block/value.
// ------------------------------------------------------------------------------------
template <int D, int C, int F>
__global__ void __launch_bounds__(kBlock)
latent_bwd_kernel(const float* __restrict__ coords, int64_t n, const float* __restrict__ grad_out,
                  const float* __restrict__ zsave, const __grid_constant__ LevelParams lp,
                  const float* __restrict__ A, int per_level, uint32_t skip_mask, uint32_t level_mask,
                  float* __restrict__ grad_latents, float* __restrict__ grad_A, float* __restrict__ grad_shift) {
    extern __shared__ float s_mem[];  // A [nA*C*F] | gA [L*C*F] | gS [L*F]
    const int L = lp.num_lods;
    const int nA = per_level ? L : 1;
    float* s_A = s_mem;
    float* s_gA = s_A + nA * C * F;
    float* s_gS = s_gA + L * C * F;
    for (int e = threadIdx.x; e < nA * C * F; e += kBlock) s_A[e] = A[e];
    for (int e = threadIdx.x; e < L * C * F + L * F; e += kBlock) s_gA[e] = 0.0f;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool live = i < n;
    const bool want_dec = (grad_A != nullptr) || (grad_shift != nullptr);
    const int lane = threadIdx.x & 31;
    double t[D];
    if (live) load_unit_coords<D>(coords, i, t);
    const float* g_row = grad_out + i * (int64_t)L * F;
    const float* z_row = zsave ? zsave + i * (int64_t)L * C : nullptr;
    constexpr int NC = 1 << D;
    const bool vec_g = (F == 1) || ((L * F) % (F >= 4 ? 4 : F) == 0);
    const bool vec_z = (C == 1) || ((L * C) % (C >= 4 ? 4 : C) == 0);
#pragma unroll 2
    for (int l = 0; l < L; ++l) {
        if (!((level_mask >> l) & 1u)) continue;  // not part of this launch (level-chunked backward)
        // levels accumulated in shared memory by coarse_bwd_kernel: only the decoder gradients remain here
        const bool scatter = !((skip_mask >> l) & 1u);
        if (!scatter && !want_dec) continue;
        float g[F];
#pragma unroll
        for (int j = 0; j < F; ++j) g[j] = 0.0f;
        if (live) {
            if (vec_g) {
                load_row<F>(g_row + l * F, g);
            } else {
#pragma unroll
                for (int j = 0; j < F; ++j) g[j] = __ldg(g_row + l * F + j);
            }
        }
        if (live && scatter) {
            Corners<D> c;
            corners<D>(t, lp, l, c);
            const int la = per_level ? l : 0;
            float gz[C];
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                float acc = 0.0f;
#pragma unroll
                for (int j = 0; j < F; ++j) acc = __fmaf_rn(g[j], s_A[(la * C + ch) * F + j], acc);
                gz[ch] = acc;
            }
            float* base = grad_latents + (int64_t)lp.first[l] * C;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                float gv[C];
#pragma unroll
                for (int ch = 0; ch < C; ++ch) gv[ch] = __fmul_rn(gz[ch], c.w[k]);
                red_add_row<C>(base + (int64_t)c.idx[k] * C, gv);
            }
        }
        if (want_dec) {  // uniform across the block
            float z[C];
#pragma unroll
            for (int ch = 0; ch < C; ++ch) z[ch] = 0.0f;
            if (live && z_row) {
                if (vec_z) {
                    load_row<C>(z_row + l * C, z);
                } else {
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) z[ch] = __ldg(z_row + l * C + ch);
                }
            }
#pragma unroll
            for (int j = 0; j < F; ++j) {
                const float sg = warp_sum(g[j]);
                if (lane == 0) atomicAdd(&s_gS[l * F + j], sg);
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const float sa = warp_sum(z[ch] * g[j]);
                    if (lane == 0) atomicAdd(&s_gA[(l * C + ch) * F + j], sa);
                }
            }
        }
    }
    if (want_dec) {
        __syncthreads();
        if (grad_A)
            for (int e = threadIdx.x; e < L * C * F; e += kBlock) red_add(grad_A + e, s_gA[e]);
        if (grad_shift)
            for (int e = threadIdx.x; e < L * F; e += kBlock) red_add(grad_shift + e, s_gS[e]);
    }
}

// ------------------------------------------------------------------------------------
// corner dump
// ------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kBlock)
corners_kernel(const float* __restrict__ coords, int64_t n, const __grid_constant__ LevelParams lp,
               int32_t* __restrict__ idx, float* __restrict__ w) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    double t[D];
    load_unit_coords<D>(coords, i, t);
    constexpr int NC = 1 << D;
    for (int l = 0; l < lp.num_lods; ++l) {
        Corners<D> c;
        corners<D>(t, lp, l, c);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            idx[(i * lp.num_lods + l) * NC + k] = c.idx[k];
            w[(i * lp.num_lods + l) * NC + k] = c.w[k];
        }
    }
}

}  // namespace shacira
