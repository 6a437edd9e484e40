// entropy_kernels.cuh -- fused factorized-density bit-rate estimate (forward + backward in one
// pass) and the symbol/histogram kernels behind LatentGrid.size().
//
// Reference: LatentGrid.ent_loss  wisp/models/grids/latent_grid.py:122-136
//            Bitparm/BitEstimator  wisp/models/prob_models/bit_estimator.py:9-65
//            LatentGrid.size       wisp/models/grids/latent_grid.py:138-153
#pragma once
#include "common.cuh"

namespace shacira {

constexpr int kEntBlock = 256;
constexpr int kMaxEntC = 16;

struct LevelBounds {
    int32_t first[SHACIRA_MAX_LEVELS + 1];  // first row of each level, then total rows
    int32_t num_lods;
};

// tanh / sigmoid through ex2.approx (__expf, ~2 ulp) and the approximate divide: absolute error ~2e-7, far
// inside the 1e-5 (bits) / 1e-4 (gradients) gates, at a fraction of the instruction count of tanhf / expf
// (the accurate versions made this kernel instruction-bound: 750 instructions per table entry).
__device__ __forceinline__ float fast_sigmoid(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float fast_tanh(float u) {
    const float a = fminf(fabsf(u), 15.0f);
    const float t = 1.0f - __fdividef(2.0f, __expf(2.0f * a) + 1.0f);
    return copysignf(t, u);
}

// One evaluation of the CDF chain at x, keeping what the reverse sweep needs.
struct CdfTrace {
    float xin[3];  // input of non-final layer k
    float th[3];   // tanh(u_k)
    float xf;      // input of the final layer
    float F;       // sigmoid output
};

__device__ __forceinline__ float cdf_forward(float x, int m, const float* sp, const float* b, const float* ta,
                                             int C, int ch, CdfTrace& tr) {
    // non-final layers f1..fm: u = x*softplus(h) + b ; x = u + tanh(u)*tanh(a)   (bit_estimator.py:43-44)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k < m) {
            tr.xin[k] = x;
            const float u = fmaf(x, sp[k * C + ch], b[k * C + ch]);
            const float th = fast_tanh(u);
            tr.th[k] = th;
            x = fmaf(th, ta[k * C + ch], u);
        }
    }
    // final layer f4: sigmoid(x*softplus(h) + b)                                  (bit_estimator.py:41)
    tr.xf = x;
    const float v = fmaf(x, sp[3 * C + ch], b[3 * C + ch]);
    tr.F = fast_sigmoid(v);
    return tr.F;
}

// Reverse sweep: given G = d(loss)/dF, accumulate d/d{softplus(h), b, tanh(a)} per layer and
// return d(loss)/dx.
__device__ __forceinline__ float cdf_backward(float G, int m, const float* sp, const float* ta, int C, int ch,
                                              const CdfTrace& tr, float (&d_sp)[4], float (&d_b)[4],
                                              float (&d_ta)[4]) {
    float gv = G * tr.F * (1.0f - tr.F);
    d_sp[3] += gv * tr.xf;
    d_b[3] += gv;
    float gx = gv * sp[3 * C + ch];
#pragma unroll
    for (int k = 2; k >= 0; --k) {
        if (k < m) {
            const float th = tr.th[k];
            d_ta[k] += gx * th;
            const float gu = gx * fmaf(1.0f - th * th, ta[k * C + ch], 1.0f);
            d_sp[k] += gu * tr.xin[k];
            d_b[k] += gu;
            gx = gu * sp[k * C + ch];
        }
    }
    return gx;
}

// params [4][3][C] = {f1,f2,f3,f4} x {h,b,a}. total = rows*C elements, element e -> (row e/C, ch e%C).
__global__ void __launch_bounds__(kEntBlock)
entropy_kernel(const float* __restrict__ latents, const float* __restrict__ noise, int64_t total, int C,
               const float* __restrict__ params, int num_layers, const __grid_constant__ LevelBounds lb,
               double* __restrict__ bits, float* __restrict__ grad_latents, float* __restrict__ grad_params,
               float* __restrict__ partials, unsigned* __restrict__ ticket, unsigned long long rng_seed,
               unsigned long long* __restrict__ rng_step) {
    __shared__ float s_sp[4 * kMaxEntC], s_b[4 * kMaxEntC], s_ta[4 * kMaxEntC];
    __shared__ float s_dsp[4 * kMaxEntC], s_dta[4 * kMaxEntC];  // chain-rule factors
    __shared__ float s_lvl[SHACIRA_MAX_LEVELS];
    __shared__ float s_acc[3 * 4 * kMaxEntC];
    __shared__ double s_total;
    const int tid = threadIdx.x;
    if (tid < 4 * C) {
        const int k = tid / C, ch = tid % C;
        const float h = params[(k * 3 + 0) * C + ch];
        const float a = params[(k * 3 + 2) * C + ch];
        // F.softplus(h): beta 1, threshold 20
        s_sp[tid] = (h > 20.0f) ? h : log1pf(expf(h));
        s_dsp[tid] = 1.0f / (1.0f + expf(-h));  // d softplus / dh
        s_b[tid] = params[(k * 3 + 1) * C + ch];
        const float t = (k < 3) ? tanhf(a) : 0.0f;
        s_ta[tid] = t;
        s_dta[tid] = 1.0f - t * t;
    }
    if (tid < SHACIRA_MAX_LEVELS) s_lvl[tid] = 0.0f;
    for (int e = tid; e < 3 * 4 * kMaxEntC; e += kEntBlock) s_acc[e] = 0.0f;
    if (tid == 0) s_total = 0.0;
    __syncthreads();

    const int m = min(num_layers, 4) - 1;  // non-final layers in use (bit_estimator.py:58-64)
    const int lane = tid & 31;
    const float inv_ln2 = 1.0f / 0.6931471805599453f;
    float d_sp[4] = {0, 0, 0, 0}, d_b[4] = {0, 0, 0, 0}, d_ta[4] = {0, 0, 0, 0};
    float my_bits = 0.0f;
    const int64_t stride = (int64_t)gridDim.x * kEntBlock;  // multiple of C (C divides 256)
    const int ch = (int)(((int64_t)blockIdx.x * kEntBlock + tid) % C);
    const int64_t rounds = (total + stride - 1) / stride;
    // in-kernel training noise: U(-0.5, 0.5) from a counter-based hash of (element, step, seed); the step counter
    // lives on the device and is advanced by the last CTA, so a captured graph draws fresh noise on every replay
    const bool rng = rng_step != nullptr;
    const bool train = rng || noise != nullptr;
    uint32_t rng_base = 0u;
    if (rng) {
        const unsigned long long st = *rng_step;
        rng_base = (uint32_t)st * 0x9E3779B9u + (uint32_t)(st >> 32) * 0x7F4A7C15u + (uint32_t)rng_seed * 0x85EBCA6Bu +
                   (uint32_t)(rng_seed >> 32) * 0xC2B2AE35u;
    }
    for (int64_t r = 0; r < rounds; ++r) {
        const int64_t e = r * stride + (int64_t)blockIdx.x * kEntBlock + tid;
        const bool live = e < total;
        float bval = 0.0f;
        int lvl = 0;
        if (live) {
            const float w = __ldg(latents + e);
            float x;
            if (rng) {
                uint32_t h = (uint32_t)e + rng_base;      // lowbias32 finaliser, two multiply-xorshift rounds
                h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
                h += (uint32_t)(e >> 32) * 0x9E3779B9u;
                x = w + ((float)(h >> 8) * 5.9604644775390625e-08f - 0.5f);   // 24-bit uniform in [0, 1) - 0.5
            } else {
                x = noise ? (w + __ldg(noise + e)) : rintf(w);  // latent_grid.py:132
            }
            CdfTrace up, lo;
            const float Fu = cdf_forward(x + 0.5f, m, s_sp, s_b, s_ta, C, ch, up);
            const float Fl = cdf_forward(x - 0.5f, m, s_sp, s_b, s_ta, C, ch, lo);
            const float p = Fu - Fl;
            const float raw = -logf(p + 1e-10f) * inv_ln2;  // latent_grid.py:135
            bval = fminf(fmaxf(raw, 0.0f), 50.0f);
            // d clamp: pass-through inside [0, 50] (inclusive, as torch.clamp)
            const float g_raw = (raw >= 0.0f && raw <= 50.0f) ? 1.0f : 0.0f;
            const float g_p = -g_raw * inv_ln2 / (p + 1e-10f);
            const float gx_u = cdf_backward(g_p, m, s_sp, s_ta, C, ch, up, d_sp, d_b, d_ta);
            const float gx_l = cdf_backward(-g_p, m, s_sp, s_ta, C, ch, lo, d_sp, d_b, d_ta);
            if (grad_latents) grad_latents[e] = train ? (gx_u + gx_l) : 0.0f;  // round() has zero gradient
            if (lb.num_lods > 0) {
                const int32_t row = (int32_t)(e / C);
                int a = 0, bnd = lb.num_lods;  // last level whose first row <= row
                while (bnd - a > 1) {
                    const int mid = (a + bnd) >> 1;
                    if (lb.first[mid] <= row) a = mid; else bnd = mid;
                }
                lvl = a;
            }
        }
        my_bits += bval;
        if (lb.num_lods > 0) {
            const int l0 = __shfl_sync(0xffffffffu, lvl, 0);
            const float s = warp_sum((live && lvl == l0) ? bval : 0.0f);
            if (lane == 0) atomicAdd(&s_lvl[l0], s);
            if (live && lvl != l0) atomicAdd(&s_lvl[lvl], bval);
        }
    }
    // block reduction: total bits (double) and parameter gradients per channel
    const float wsum = warp_sum(my_bits);
    if (lane == 0) atomicAdd(&s_total, (double)wsum);
    if (grad_params) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v0 = d_sp[k], v1 = d_b[k], v2 = d_ta[k];
            // lanes l and l^o share a channel when o is a multiple of C
            for (int o = 16; o >= C; o >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            }
            if (lane < C) {
                const int c2 = (C <= 32) ? ((tid & ~31) + lane) % C : 0;
                atomicAdd(&s_acc[(k * 3 + 0) * kMaxEntC + c2], v0);
                atomicAdd(&s_acc[(k * 3 + 1) * kMaxEntC + c2], v1);
                atomicAdd(&s_acc[(k * 3 + 2) * kMaxEntC + c2], v2);
            }
        }
    }
    __syncthreads();
    // Block partials go to scratch; the LAST block to arrive (ticket) sums them in a fixed order and writes the
    // outputs. No same-address global atomics (they serialise in L2: ~30 per block x 1000+ blocks measured at
    // ~10 us) and no zero-fill of the outputs.
    //   partial row layout: [0] total | [1 .. 1+L) per level | [1+L .. 1+L+12*C) parameter gradients
    const int L = lb.num_lods;
    const int P = 1 + L + 12 * C;
    float* mine = partials + (size_t)blockIdx.x * P;
    if (tid == 0) mine[0] = (float)s_total;
    if (tid < L) mine[1 + tid] = s_lvl[tid];
    if (tid < 4 * C) {
        const int k = tid / C, c2 = tid % C;
        // chain to the raw parameters: h through softplus, a through tanh
        mine[1 + L + (k * 3 + 0) * C + c2] = s_acc[(k * 3 + 0) * kMaxEntC + c2] * s_dsp[tid];
        mine[1 + L + (k * 3 + 1) * C + c2] = s_acc[(k * 3 + 1) * kMaxEntC + c2];
        mine[1 + L + (k * 3 + 2) * C + c2] = (k < 3) ? s_acc[(k * 3 + 2) * kMaxEntC + c2] * s_dta[tid] : 0.0f;
    }
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // one warp per output value, lanes stride over the blocks (independent loads), fixed summation order
    for (int v = tid >> 5; v < P; v += kEntBlock / 32) {
        double sum = 0.0;
        for (unsigned b = lane; b < gridDim.x; b += 32) sum += (double)partials[(size_t)b * P + v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) {
            if (v <= L) bits[v] = sum;
            else if (grad_params) grad_params[v - 1 - L] = (float)sum;
        }
    }
    if (tid == 0) {
        *ticket = 0u;  // ready for the next launch on this scratch
        if (rng) *rng_step += 1ull;  // every CTA has read it (this is the last one to arrive)
    }
}

// ---- validation mode (x = round(w)) through a per-integer table -------------------------------------------------------
// The NeRF trainer evaluates the bit-rate loss on round(w) every step (multiview_trainer.py:110) and
// LatentGrid.size()-style reports do the same: the CDF chain then depends on (channel, integer) only. Every CTA builds the
// table of the 256 integers around 0 in its prologue (one thread per integer: value and the 12 parameter-gradient terms,
// cdf_forward / cdf_backward as in the per-element kernel, so the numbers are the same) and the table's elements then cost
// 13 shared-memory loads + 13 adds each instead of ~300 dependent instructions with two transcendental chains. Integers
// outside [-128, 127] and NaN take the per-element evaluation. Contiguous element ranges per CTA, so a thread walks the
// levels monotonically (no per-element level search). Partial-row layout, ticket and final reduction as entropy_kernel.
constexpr int kLutN = 256, kLutK = 13;   // integers per channel; bits | d softplus(h) x4 | d b x4 | d tanh(a) x4
__global__ void __launch_bounds__(kEntBlock)   // (capped at 64 registers for 4 CTAs per SM it spills: 56 vs 34 us)
entropy_val_lut_kernel(const float* __restrict__ latents, int64_t total, int C, const float* __restrict__ params,
                       int num_layers, const __grid_constant__ LevelBounds lb, double* __restrict__ bits,
                       float* __restrict__ grad_latents, float* __restrict__ grad_params, float* __restrict__ partials,
                       unsigned* __restrict__ ticket, int64_t per_block) {
    extern __shared__ float s_lut[];   // [C][kLutK][kLutN]
    __shared__ float s_sp[4 * kMaxEntC], s_b[4 * kMaxEntC], s_ta[4 * kMaxEntC];
    __shared__ float s_dsp[4 * kMaxEntC], s_dta[4 * kMaxEntC];
    __shared__ float s_lvl[SHACIRA_MAX_LEVELS];
    __shared__ float s_acc[3 * 4 * kMaxEntC];
    __shared__ double s_total;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 4 * C) {
        const int k = tid / C, ch = tid % C;
        const float h = params[(k * 3 + 0) * C + ch];
        const float a = params[(k * 3 + 2) * C + ch];
        s_sp[tid] = (h > 20.0f) ? h : log1pf(expf(h));
        s_dsp[tid] = 1.0f / (1.0f + expf(-h));
        s_b[tid] = params[(k * 3 + 1) * C + ch];
        const float t = (k < 3) ? tanhf(a) : 0.0f;
        s_ta[tid] = t;
        s_dta[tid] = 1.0f - t * t;
    }
    if (tid < SHACIRA_MAX_LEVELS) s_lvl[tid] = 0.0f;
    for (int e = tid; e < 3 * 4 * kMaxEntC; e += kEntBlock) s_acc[e] = 0.0f;
    if (tid == 0) s_total = 0.0;
    __syncthreads();
    const int m = min(num_layers, 4) - 1;
    const float inv_ln2 = 1.0f / 0.6931471805599453f;
    // one element's value and parameter-gradient terms at x (the per-element kernel's formulas)
    auto eval = [&](float x, int ch, float (&out)[kLutK]) {
        CdfTrace up, lo;
        const float Fu = cdf_forward(x + 0.5f, m, s_sp, s_b, s_ta, C, ch, up);
        const float Fl = cdf_forward(x - 0.5f, m, s_sp, s_b, s_ta, C, ch, lo);
        const float p = Fu - Fl;
        const float raw = -logf(p + 1e-10f) * inv_ln2;
        const float g_raw = (raw >= 0.0f && raw <= 50.0f) ? 1.0f : 0.0f;
        const float g_p = -g_raw * inv_ln2 / (p + 1e-10f);
        float d_sp[4] = {0, 0, 0, 0}, d_b[4] = {0, 0, 0, 0}, d_ta[4] = {0, 0, 0, 0};
        cdf_backward(g_p, m, s_sp, s_ta, C, ch, up, d_sp, d_b, d_ta);
        cdf_backward(-g_p, m, s_sp, s_ta, C, ch, lo, d_sp, d_b, d_ta);
        out[0] = fminf(fmaxf(raw, 0.0f), 50.0f);
#pragma unroll
        for (int k = 0; k < 4; ++k) { out[1 + k] = d_sp[k]; out[5 + k] = d_b[k]; out[9 + k] = d_ta[k]; }
    };
    for (int ch = 0; ch < C; ++ch) {   // thread tid <-> integer tid - 128
        float v[kLutK];
        eval((float)(tid - kLutN / 2), ch, v);
#pragma unroll
        for (int k = 0; k < kLutK; ++k) s_lut[(ch * kLutK + k) * kLutN + tid] = v[k];
    }
    __syncthreads();

    // terms that can be non-zero with m non-final layers in use: bits, the final layer's (softplus, b), layers j < m
    unsigned used = 1u | (1u << 4) | (1u << 8);
    for (int j = 0; j < m; ++j) used |= (1u << (1 + j)) | (1u << (5 + j)) | (1u << (9 + j));
    const int64_t base = (int64_t)blockIdx.x * per_block, end = min(total, base + per_block);
    const int ch = (int)((base + tid) % C);   // per_block and 256 are multiples of C: fixed per thread
    const float* lut = s_lut + (size_t)ch * kLutK * kLutN;
    float acc[kLutK];
#pragma unroll
    for (int k = 0; k < kLutK; ++k) acc[k] = 0.0f;
    float lvl_bits = 0.0f;
    int lvl = 0;
    int64_t next_first = (lb.num_lods > 0) ? (int64_t)lb.first[1] : (int64_t)1 << 62;
    // this warp's per-level bits so far go to shared memory: one add per warp for lane 0's level, stragglers (a warp spans
    // at most two levels) add their own
    auto flush_levels = [&]() {
        const int l0 = __shfl_sync(0xffffffffu, lvl, 0);
        const float sum = warp_sum(lvl == l0 ? lvl_bits : 0.0f);
        if (lane == 0 && sum != 0.0f) atomicAdd(&s_lvl[l0], sum);
        if (lvl != l0 && lvl_bits != 0.0f) atomicAdd(&s_lvl[lvl], lvl_bits);
        lvl_bits = 0.0f;
    };
    constexpr int U = 4;   // elements per thread in flight (the loop is otherwise bound by one load latency per element)
    // warp-uniform trip count (the level flush votes across the warp): the warp's first lane decides, lanes past `end` idle
    for (int64_t ew = base + (tid & ~31); ew < end; ew += (int64_t)U * kEntBlock) {
        const int64_t e0 = ew + lane;
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = e0 + (int64_t)u * kEntBlock;
            w[u] = (e < end) ? __ldg(latents + e) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = e0 + (int64_t)u * kEntBlock;
            const bool live = e < end;
            const float xq = rintf(w[u]);
            float v[kLutK];
            if (xq >= -128.0f && xq <= 127.0f) {
                const int q = (int)xq + kLutN / 2;
#pragma unroll
                for (int k = 0; k < kLutK; ++k) v[k] = (live && ((used >> k) & 1u)) ? lut[k * kLutN + q] : 0.0f;   // (uniform)
            } else {
                eval(xq, ch, v);   // far-out integers, Inf, NaN: the per-element evaluation
            }
            if (live && grad_latents) grad_latents[e] = 0.0f;   // round() has zero gradient
            if (lb.num_lods > 0) {
                const int64_t row = e / C;
                if (__any_sync(0xffffffffu, live && row >= next_first)) {
                    flush_levels();
                    while (live && lvl + 1 < lb.num_lods && row >= (int64_t)lb.first[lvl + 1]) ++lvl;
                    next_first = (lvl + 1 < lb.num_lods) ? (int64_t)lb.first[lvl + 1] : (int64_t)1 << 62;
                }
                lvl_bits += v[0];
            }
#pragma unroll
            for (int k = 0; k < kLutK; ++k) acc[k] += v[k];
        }
    }
    if (lb.num_lods > 0) flush_levels();
    const float wsum = warp_sum(acc[0]);
    if (lane == 0) atomicAdd(&s_total, (double)wsum);
    if (grad_params) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v0 = acc[1 + k], v1 = acc[5 + k], v2 = acc[9 + k];
            for (int o = 16; o >= C; o >>= 1) {   // lanes l and l^o share a channel when o is a multiple of C
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            }
            if (lane < C) {
                const int c2 = (int)((base + (tid & ~31) + lane) % C);
                atomicAdd(&s_acc[(k * 3 + 0) * kMaxEntC + c2], v0);
                atomicAdd(&s_acc[(k * 3 + 1) * kMaxEntC + c2], v1);
                atomicAdd(&s_acc[(k * 3 + 2) * kMaxEntC + c2], v2);
            }
        }
    }
    __syncthreads();
    const int L = lb.num_lods;
    const int P = 1 + L + 12 * C;
    float* mine = partials + (size_t)blockIdx.x * P;
    if (tid == 0) mine[0] = (float)s_total;
    if (tid < L) mine[1 + tid] = s_lvl[tid];
    if (tid < 4 * C) {
        const int k = tid / C, c2 = tid % C;
        mine[1 + L + (k * 3 + 0) * C + c2] = s_acc[(k * 3 + 0) * kMaxEntC + c2] * s_dsp[tid];
        mine[1 + L + (k * 3 + 1) * C + c2] = s_acc[(k * 3 + 1) * kMaxEntC + c2];
        mine[1 + L + (k * 3 + 2) * C + c2] = (k < 3) ? s_acc[(k * 3 + 2) * kMaxEntC + c2] * s_dta[tid] : 0.0f;
    }
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int v = tid >> 5; v < P; v += kEntBlock / 32) {
        double sum = 0.0;
        for (unsigned b = lane; b < gridDim.x; b += 32) sum += (double)partials[(size_t)b * P + v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) {
            if (v <= L) bits[v] = sum;
            else if (grad_params) grad_params[v - 1 - L] = (float)sum;
        }
    }
    if (tid == 0) *ticket = 0u;
}

// ---- fused Adam over one tensor ---------------------------------------------------
// torch.optim.Adam's multi-tensor kernel walks a single tensor in 64 K-element chunks -- 6 CTAs for the
// 375 k-row latent table of the image fit (39 us measured). One thread per 4 elements here (~3 us), same update:
//   g += wd * p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr / (1-b1^t) * m / (sqrt(v / (1-b2^t)) + eps)
// `step` is a device counter (float, as torch keeps it) incremented by the kernel: CUDA-graph capturable.
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                 float* __restrict__ step, int zero_grad, float* __restrict__ g_mut) {
    const float t = *step + 1.0f;
    const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const int64_t i4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i4 + 3 < n && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                        reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0) {
        float4 pp = *reinterpret_cast<float4*>(p + i4), mm = *reinterpret_cast<float4*>(m + i4);
        float4 vv = *reinterpret_cast<float4*>(v + i4);
        const float4 gg = *reinterpret_cast<const float4*>(g + i4);
        float* P = &pp.x; float* M = &mm.x; float* V = &vv.x; const float* G = &gg.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = fmaf(weight_decay, P[k], G[k]);
            M[k] = fmaf(beta1, M[k], (1.0f - beta1) * gk);
            V[k] = fmaf(beta2, V[k], (1.0f - beta2) * gk * gk);
            P[k] -= step_size * M[k] / (sqrtf(V[k]) * inv_sqrt_bc2 + eps);
        }
        *reinterpret_cast<float4*>(p + i4) = pp;
        *reinterpret_cast<float4*>(m + i4) = mm;
        *reinterpret_cast<float4*>(v + i4) = vv;
        if (zero_grad) *reinterpret_cast<float4*>(g_mut + i4) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int64_t i = i4; i < min(n, i4 + 4); ++i) {
            const float gk = fmaf(weight_decay, p[i], g[i]);
            const float mk = fmaf(beta1, m[i], (1.0f - beta1) * gk);
            const float vk = fmaf(beta2, v[i], (1.0f - beta2) * gk * gk);
            m[i] = mk;
            v[i] = vk;
            p[i] -= step_size * mk / (sqrtf(vk) * inv_sqrt_bc2 + eps);
            if (zero_grad) g_mut[i] = 0.0f;
        }
    }
}
__global__ void adam_advance_kernel(float* step) { *step += 1.0f; }

// grad = g + (*scale2 * mul2) * g2: the bit-rate gradient rides on the grid gradient (no pass of its own)
__global__ void __launch_bounds__(256)
adam_step_sum_kernel(float* __restrict__ p, const float* __restrict__ g, const float* __restrict__ gmul,
                     const float* __restrict__ g2,
                     const float* __restrict__ scale2, float mul2, float* __restrict__ m, float* __restrict__ v,
                     int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                     const float* __restrict__ step, int zero_grad) {
    const float t = *step + 1.0f;
    const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const float s2 = g2 ? (scale2 ? *scale2 * mul2 : mul2) : 0.0f;
    const int64_t i4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    const bool vec = i4 + 3 < n && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                                     reinterpret_cast<uintptr_t>(g2) | reinterpret_cast<uintptr_t>(m) |
                                     reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(gmul)) & 15) == 0;
    if (vec) {
        float4 pp = *reinterpret_cast<float4*>(p + i4), mm = *reinterpret_cast<float4*>(m + i4);
        float4 vv = *reinterpret_cast<float4*>(v + i4);
        float4 gg = *reinterpret_cast<const float4*>(g + i4);
        if (gmul) {   // chain rule of a table-side quantiser (SGA): grid gradient times d w_hat / d w
            const float4 qq = *reinterpret_cast<const float4*>(gmul + i4);
            gg.x *= qq.x; gg.y *= qq.y; gg.z *= qq.z; gg.w *= qq.w;
        }
        if (g2) {
            const float4 hh = *reinterpret_cast<const float4*>(g2 + i4);
            gg.x = fmaf(s2, hh.x, gg.x); gg.y = fmaf(s2, hh.y, gg.y); gg.z = fmaf(s2, hh.z, gg.z); gg.w = fmaf(s2, hh.w, gg.w);
        }
        float* P = &pp.x; float* M = &mm.x; float* V = &vv.x; const float* G = &gg.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = fmaf(weight_decay, P[k], G[k]);
            M[k] = fmaf(beta1, M[k], (1.0f - beta1) * gk);
            V[k] = fmaf(beta2, V[k], (1.0f - beta2) * gk * gk);
            P[k] -= step_size * M[k] / (sqrtf(V[k]) * inv_sqrt_bc2 + eps);
        }
        *reinterpret_cast<float4*>(p + i4) = pp;
        *reinterpret_cast<float4*>(m + i4) = mm;
        *reinterpret_cast<float4*>(v + i4) = vv;
        if (zero_grad) *reinterpret_cast<float4*>(const_cast<float*>(g) + i4) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int64_t i = i4; i < min(n, i4 + 4); ++i) {
            const float g0 = gmul ? g[i] * gmul[i] : g[i];
            const float gi = g2 ? fmaf(s2, g2[i], g0) : g0;
            if (zero_grad) const_cast<float*>(g)[i] = 0.0f;
            const float gk = fmaf(weight_decay, p[i], gi);
            const float mk = fmaf(beta1, m[i], (1.0f - beta1) * gk);
            const float vk = fmaf(beta2, v[i], (1.0f - beta2) * gk * gk);
            m[i] = mk;
            v[i] = vk;
            p[i] -= step_size * mk / (sqrtf(vk) * inv_sqrt_bc2 + eps);
        }
    }
}

// Adam over many small tensors, one CTA (see shacira_multi_adam_step in the header for the gradient formula).
struct AdamSegs {
    shacira_adam_seg_t seg[SHACIRA_MAX_ADAM_SEGS];
    int32_t num;
    int32_t pad[3];
};
// One CTA per segment (a single CTA walking ~20 tiny tensors pays ~1 us of load latency per tensor: 19 us measured).
// The last CTA to finish (ticket) advances the step counters -- every CTA has read them by then -- and leaves the
// caller's ticket at 0 for the next launch.
__global__ void __launch_bounds__(128)
multi_adam_kernel(const __grid_constant__ AdamSegs S, float beta1, float beta2, float eps, float* __restrict__ step,
                  float* __restrict__ extra_step, const float* __restrict__ scale, const float* __restrict__ div,
                  float* __restrict__ A_out, int C, int F, unsigned* __restrict__ ticket) {
    const float t = *step + 1.0f;
    const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    const shacira_adam_seg_t& sg = S.seg[blockIdx.x];
    const float gs = sg.grad_mul * (sg.grad_scale ? *sg.grad_scale : 1.0f);
    const float step_size = sg.lr / bc1;
    for (int i = threadIdx.x; i < sg.n; i += 128) {
        float g = 0.0f;
        for (int r = 0; r < sg.grad_rows; ++r) g += sg.grad[(size_t)r * sg.grad_row_stride + i];
        if (sg.zero_grad != 0.0f)
            for (int r = 0; r < sg.grad_rows; ++r) const_cast<float*>(sg.grad)[(size_t)r * sg.grad_row_stride + i] = 0.0f;
        g *= gs;
        if (sg.grad_div) g /= sg.grad_div[i / sg.div_group];
        const float p = sg.param[i];
        const float gk = fmaf(sg.weight_decay, p, g);
        const float mk = fmaf(beta1, sg.exp_avg[i], (1.0f - beta1) * gk);
        const float vk = fmaf(beta2, sg.exp_avg_sq[i], (1.0f - beta2) * gk * gk);
        sg.exp_avg[i] = mk;
        sg.exp_avg_sq[i] = vk;
        sg.param[i] = p - step_size * mk / (sqrtf(vk) * inv_sqrt_bc2 + eps);
    }
    __syncthreads();  // the block's updated values are visible to the block
    // the CTA that owns the latent decoder's scale refreshes A = scale / div for the next step's kernels
    if (A_out && sg.param == scale)
        for (int e = threadIdx.x; e < C * F; e += 128) A_out[e] = scale[e] / div[e / F];
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *step = t;
            if (extra_step) *extra_step += 1.0f;
            *ticket = 0u;
        }
    }
}

// ---- symbols and histogram ---------------------------------------------------------------
__global__ void init_minmax_kernel(int32_t* minmax, int C) {
    const int c = threadIdx.x;
    if (c < C) {
        minmax[2 * c + 0] = INT32_MAX;
        minmax[2 * c + 1] = INT32_MIN;
    }
}

__global__ void __launch_bounds__(256)
quantize_symbols_kernel(const float* __restrict__ latents, int64_t total, int C, int16_t* __restrict__ symbols,
                        int32_t* __restrict__ minmax) {
    __shared__ int32_t s_min[kMaxEntC], s_max[kMaxEntC];
    if (threadIdx.x < C) {
        s_min[threadIdx.x] = INT32_MAX;
        s_max[threadIdx.x] = INT32_MIN;
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * 256;
    const int ch = (int)(((int64_t)blockIdx.x * 256 + threadIdx.x) % C);
    int32_t lo = INT32_MAX, hi = INT32_MIN;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += stride) {
        const int32_t q = __float2int_rn(__ldg(latents + e));  // torch.round: half to even
        if (symbols) symbols[e] = (int16_t)q;
        lo = min(lo, q);
        hi = max(hi, q);
    }
    if (lo <= hi) {
        atomicMin(&s_min[ch], lo);
        atomicMax(&s_max[ch], hi);
    }
    __syncthreads();
    if (threadIdx.x < C && s_min[threadIdx.x] <= s_max[threadIdx.x]) {
        atomicMin(&minmax[2 * threadIdx.x + 0], s_min[threadIdx.x]);
        atomicMax(&minmax[2 * threadIdx.x + 1], s_max[threadIdx.x]);
    }
}

constexpr int kHistSmemBins = 8192;

__global__ void __launch_bounds__(256)
symbol_histogram_kernel(const float* __restrict__ latents, int64_t total, int C, const int32_t* __restrict__ lo,
                        int num_bins, unsigned long long* __restrict__ counts) {
    __shared__ uint32_t s_hist[kHistSmemBins];
    const bool use_smem = (int64_t)C * num_bins <= kHistSmemBins;
    if (use_smem) {
        for (int e = threadIdx.x; e < C * num_bins; e += 256) s_hist[e] = 0u;
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * 256;
    const int ch = (int)(((int64_t)blockIdx.x * 256 + threadIdx.x) % C);
    const int32_t off = lo[ch];
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += stride) {
        const int32_t bin = __float2int_rn(__ldg(latents + e)) - off;
        if (bin < 0 || bin >= num_bins) continue;  // caller sizes the bins from the min/max pass
        if (use_smem) atomicAdd(&s_hist[ch * num_bins + bin], 1u);
        else atomicAdd(&counts[(int64_t)ch * num_bins + bin], 1ull);
    }
    if (use_smem) {
        __syncthreads();
        for (int e = threadIdx.x; e < C * num_bins; e += 256)
            if (s_hist[e]) atomicAdd(&counts[e], (unsigned long long)s_hist[e]);
    }
}

}  // namespace shacira
