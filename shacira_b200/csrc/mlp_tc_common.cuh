// mlp_tc_common.cuh -- warp-level building blocks of the decoder MLP on the tensor cores (mma.sync m16n8k8 TF32 with
// 3xTF32 compensation), shared by the stand-alone MLP kernel (mlp_kernels.cuh) and the fused fit kernel
// (fit_kernels.cuh). See the comment block above mlp16_tc_step_kernel for the layout trick.
#pragma once
#include "common.cuh"

namespace shacira {

constexpr int kTcStride = 24;   // floats per staged point row

// v = hi + lo with hi, lo TF32 (10 explicit mantissa bits). Round-to-nearest on the magnitude by integer add + mask:
// `cvt.rna.tf32.f32` compiles to a ~5-instruction sequence on sm_100a (cuobjdump) and the step needs ~250 splits per
// 32 points. The tensor core ignores the low 13 bits of a TF32 operand, so lo only needs the rounding add.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi)) + 0x1000u;
}
// Activations: lo is left unrounded (the tensor core truncates it to 10 mantissa bits: error 2^-21 |v|, the order of the
// dropped lo*lo term) -- one instruction less per split, ~90 splits per 16 points. (Measured and not adopted: hi left
// unmasked for the hardware to truncate, two instructions per split -- no change in either kernel: they are latency bound.)
__device__ __forceinline__ void split_tf32_act(float v, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += (ahi + alo) * (bhi + blo), small terms first
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], uint32_t bh0,
                                     uint32_t bh1, uint32_t bl0, uint32_t bl1) {
    mma_tf32(c, alo, bh0, bh1);
    mma_tf32(c, ahi, bl0, bl1);
    mma_tf32(c, ahi, bh0, bh1);
}
// the same with the two small terms in an accumulator of their own: two independent HMMA chains per output tile
// (SHACIRA_TC_SPLIT_CHAINS=0: one chain, 4 registers per tile less)
#ifndef SHACIRA_TC_SPLIT_CHAINS
#define SHACIRA_TC_SPLIT_CHAINS 0   // measured on B200: no gain in either kernel
#endif
__device__ __forceinline__ void mma3s(float (&c)[4], float (&sm)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                      uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
#if SHACIRA_TC_SPLIT_CHAINS
    mma_tf32(sm, alo, bh0, bh1);
    mma_tf32(c, ahi, bh0, bh1);
    mma_tf32(sm, ahi, bl0, bl1);
#else
    mma3(c, ahi, alo, bh0, bh1, bl0, bl1);
#endif
}
// accumulator-layout tile (c0 c1 | c2 c3 = rows g | g+8, columns 2t, 2t+1) -> A fragment under the K permutation
__device__ __forceinline__ void tile_to_a(const float (&c)[4], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
    split_tf32_act(c[0], hi[0], lo[0]);
    split_tf32_act(c[2], hi[1], lo[1]);
    split_tf32_act(c[1], hi[2], lo[2]);
    split_tf32_act(c[3], hi[3], lo[3]);
}

// out[mt][nt] (+)= in[mt][ks] x W-fragments; KS k-steps, NT n-tiles; fragment f(ks, nt) = wf[base + ks * NT + nt]
template <int KS, int NT, int MT>
__device__ __forceinline__ void tc_layer(const uint4 (*wf)[32], int base, int lane, const float (&in)[MT][2][4],
                                         float (&out)[MT][2][4]) {
    float sm[MT][NT][4];   // lo*hi + hi*lo terms: a second, independent accumulation chain per output tile
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) sm[mt][nt][r] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        uint4 f[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) f[nt] = wf[base + ks * NT + nt][lane];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            uint32_t ahi[4], alo[4];
            tile_to_a(in[mt][ks], ahi, alo);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma3s(out[mt][nt], sm[mt][nt], ahi, alo, f[nt].x, f[nt].y, f[nt].z, f[nt].w);
        }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) out[mt][nt][r] += sm[mt][nt][r];
}

// stage a [32 points x 16] activation held in accumulator layout as rows of kTcStride floats
template <int MT>
__device__ __forceinline__ void tc_stage(float (*rows)[kTcStride], int g, int t, const float (&a)[MT][2][4]) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            *reinterpret_cast<float2*>(&rows[16 * mt + g][8 * nt + 2 * t]) = make_float2(a[mt][nt][0], a[mt][nt][1]);
            *reinterpret_cast<float2*>(&rows[16 * mt + g + 8][8 * nt + 2 * t]) = make_float2(a[mt][nt][2], a[mt][nt][3]);
        }
}

// acc[nt] += A^T B over the warp's 32 points: A = rowsA[p][16] (M index = column of A), B = rowsB[p][8 * NT]
template <int NT, int SB, int MT>
__device__ __forceinline__ void tc_wgrad(const float (*rowsA)[kTcStride], const float (*rowsB)[SB], int g, int t,
                                         float (&acc)[NT][4]) {
    float sm[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r) sm[nt][r] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 2 * MT; ++ks) {
        const int p0 = 8 * ks + t, p1 = p0 + 4;
        uint32_t ahi[4], alo[4];
        split_tf32_act(rowsA[p0][g], ahi[0], alo[0]);
        split_tf32_act(rowsA[p0][g + 8], ahi[1], alo[1]);
        split_tf32_act(rowsA[p1][g], ahi[2], alo[2]);
        split_tf32_act(rowsA[p1][g + 8], ahi[3], alo[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32_act(rowsB[p0][8 * nt + g], bh0, bl0);
            split_tf32_act(rowsB[p1][8 * nt + g], bh1, bl1);
            mma3s(acc[nt], sm[nt], ahi, alo, bh0, bh1, bl0, bl1);
        }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[nt][r] += sm[nt][r];
}

// Weight-fragment table of the 16 -> 16 -> 16 -> 3 MLP: forward L1 (ks, nt) 0..3, L2 4..7, L3 (ks) 8..9; backward d2 (nt)
// 10..11, d1 (ks, nt) 12..15, feature gradient (ks, nt) 16..19. Built once per CTA by all `nthreads` threads.
enum { F_L1 = 0, F_L2 = 4, F_L3 = 8, B_D2 = 10, B_D1 = 12, B_GX = 16 };
__device__ __forceinline__ void tc_build_fragments(uint4 (*wf)[32], const float* __restrict__ W1,
                                                   const float* __restrict__ W2, const float* __restrict__ W3, int tid,
                                                   int nthreads) {
    constexpr int OUT = 3;
    for (int e = tid; e < 20 * 32; e += nthreads) {
        const int f = e >> 5, fl = e & 31, fg = fl >> 2, ft = fl & 3;
        float w0 = 0.0f, w1 = 0.0f;
        if (f < F_L3) {                      // y = W a : B[k = in][n = out]
            const float* W = f < F_L2 ? W1 : W2;
            const int q = f & 3, ks = q >> 1, nt = q & 1;
            w0 = W[(8 * nt + fg) * 16 + 8 * ks + 2 * ft];
            w1 = W[(8 * nt + fg) * 16 + 8 * ks + 2 * ft + 1];
        } else if (f < B_D2) {               // W3: outputs padded to 8
            const int ks = f - F_L3;
            if (fg < OUT) { w0 = W3[fg * 16 + 8 * ks + 2 * ft]; w1 = W3[fg * 16 + 8 * ks + 2 * ft + 1]; }
        } else if (f < B_D1) {               // d2 = dy W3 : B[k = out][n = j]
            const int nt = f - B_D2;
            if (2 * ft < OUT) w0 = W3[(2 * ft) * 16 + 8 * nt + fg];
            if (2 * ft + 1 < OUT) w1 = W3[(2 * ft + 1) * 16 + 8 * nt + fg];
        } else {                             // d_in = d_out W : B[k = out row][n = in column]
            const float* W = f < B_GX ? W2 : W1;
            const int q = (f - B_D1) & 3, ks = q >> 1, nt = q & 1;
            w0 = W[(8 * ks + 2 * ft) * 16 + 8 * nt + fg];
            w1 = W[(8 * ks + 2 * ft + 1) * 16 + 8 * nt + fg];
        }
        uint4 v;
        split_tf32(w0, v.x, v.z);
        split_tf32(w1, v.y, v.w);
        wf[f][fl] = v;
    }
}

}  // namespace shacira
