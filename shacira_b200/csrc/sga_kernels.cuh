// sga_kernels.cuh -- stochastic Gumbel annealing of the latents (SGA), the quantiser SHACIRA trains with for the
// first `decay_period` of every fit (kodak.yaml:43,51, nerf_lego.yaml:67-68; image_trainer.py:131-137).
//
// Reference: LatentDecoder.forward, basic_latent_decoder.py:183-191, with torch's RelaxedOneHotCategorical
// (ExpRelaxedCategorical.rsample: uniforms -> clamp_probs -> Gumbel -> (logits + gumbels) / temperature -> log-softmax,
// then exp): ~25 elementwise PyTorch kernels and a [T, C, 2] noise tensor per step there, one pass over the table here.
//   wf = floor(w), wc = wf + 1
//   lf = -tanh(clamp(w - wf)) / tau,  lc = -tanh(clamp(wc - w)) / tau        clamp to +-(1 - 1e-6)
//   (lf, lc) <- (lf, lc) - logsumexp(lf, lc)                                    Categorical normalises its logits
//   g_k = -log(-log(u_k)),  u_k ~ U(0, 1) clamped to [eps, 1 - eps]            eps = 2^-23
//   s = softmax(((lf + g_0) / tau, (lc + g_1) / tau))
//   w_hat = wf * s_0 + wc * s_1
// The value feeds the fused grid kernels with rounding switched off. Its derivative is table-side too:
//   diff_sampling (rsample; wf is a constant):  d w_hat / d w = s_0 s_1 [sech^2(w - wf) 1{..} + sech^2(wc - w) 1{..}] / tau^2
//   otherwise (sample() is not differentiated, floor passes gradients straight through): d w_hat / d w = s_0 + s_1
// (1{..}: torch.clamp passes the gradient only inside its bounds). The kernel stores it as `dw` and the table's Adam
// kernel multiplies the grid gradient by it (shacira_adam_step_sum_mul): no pass of its own in the backward.
// Noise: the caller's uniforms (parity runs inject the reference's torch.rand draw) or, with u == NULL, a counter-based
// hash of (element, *rng_step, seed) like the bit-rate kernel's -- fresh noise on every CUDA-graph replay.
#pragma once
#include "common.cuh"

namespace shacira {

constexpr int kSgaBlock = 256;

__device__ __forceinline__ float sga_uniform(uint32_t e, uint32_t base) {
    uint32_t h = e + base;   // lowbias32 finaliser
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return (float)(h >> 8) * 5.9604644775390625e-08f;   // 24-bit uniform in [0, 1)
}

// One element: w_hat = floor(w) p0 + ceil(w) p1 with (p0, p1) the relaxed one-hot sample, and d w_hat / d w
// (basic_latent_decoder.py:183-191). u0, u1: the two U(0, 1) draws of the element.
__device__ __forceinline__ void sga_sample(float x, float u0, float u1, float tau, int diff_sampling, float& w_hat, float& d) {
    const float inv_tau = 1.0f / tau;
    const float bound = (float)(1.0 - 1e-6);             // basic_latent_decoder.py:13 epsilon
    const float ueps = 1.1920928955078125e-07f;           // torch.finfo(float32).eps (clamp_probs)
    const float wf = floorf(x), wc = wf + 1.0f;
    const float df_raw = x - wf, dc_raw = wc - x;
    const float df = fminf(fmaxf(df_raw, -bound), bound), dc = fminf(fmaxf(dc_raw, -bound), bound);
    const float tf = tanhf(df), tc = tanhf(dc);
    // The reference normalises the two logits (Categorical: logits - logsumexp), adds the Gumbels, divides by tau and
    // takes a softmax. Both normalisations cancel in the only quantity the softmax of TWO entries depends on, the
    // difference of its arguments:  s0 - s1 = ((tc - tf) / tau + g0 - g1) / tau,  with
    // g0 - g1 = log(log(u1) / log(u0)).  p0 = sigmoid(s0 - s1), p1 = sigmoid(s1 - s0): 2 tanh + 3 log + 2 exp instead of
    // 2 tanh + 6 log + 4 exp, equal to the reference's chain to float rounding (golden vectors: tests/test_sga_gpu.py).
    u0 = fminf(fmaxf(u0, ueps), 1.0f - ueps);
    u1 = fminf(fmaxf(u1, ueps), 1.0f - ueps);
    const float dg = logf(logf(u1) / logf(u0));          // g0 - g1 (both logs are negative: the ratio is positive)
    const float ds = ((tc - tf) / tau + dg) / tau;         // s0 - s1
    const float e = expf(-fabsf(ds));                      // stable two-way softmax
    const float big = 1.0f / (1.0f + e), small = e * big;
    const float p0 = ds >= 0.0f ? big : small, p1 = ds >= 0.0f ? small : big;
    w_hat = __fadd_rn(__fmul_rn(wf, p0), __fmul_rn(wc, p1));
    if (diff_sampling) {
        const float in_f = (df_raw >= -bound && df_raw <= bound) ? (1.0f - tf * tf) : 0.0f;
        const float in_c = (dc_raw >= -bound && dc_raw <= bound) ? (1.0f - tc * tc) : 0.0f;
        d = p0 * p1 * (in_f + in_c) * inv_tau * inv_tau;
    } else {
        d = p0 + p1;
    }
}
__device__ __forceinline__ uint32_t sga_rng_base(unsigned long long st, unsigned long long rng_seed) {
    return (uint32_t)st * 0x9E3779B9u + (uint32_t)(st >> 32) * 0x7F4A7C15u + (uint32_t)rng_seed * 0x85EBCA6Bu +
           (uint32_t)(rng_seed >> 32) * 0xC2B2AE35u + 0x68E31DA4u;
}

__global__ void __launch_bounds__(kSgaBlock)
sga_quantize_kernel(const float* __restrict__ w, const float* __restrict__ u, int64_t n,
                    const float* __restrict__ temperature, int diff_sampling, unsigned long long rng_seed,
                    const unsigned long long* __restrict__ rng_step, float* __restrict__ w_hat, float* __restrict__ dw) {
    const float tau = __ldg(temperature);
    const uint32_t base = u ? 0u : sga_rng_base(rng_step ? *rng_step : 0ull, rng_seed);
    for (int64_t i = (int64_t)blockIdx.x * kSgaBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kSgaBlock) {
        const float u0 = u ? u[2 * i] : sga_uniform((uint32_t)(2 * i), base);
        const float u1 = u ? u[2 * i + 1] : sga_uniform((uint32_t)(2 * i + 1), base);
        float wh, d;
        sga_sample(w[i], u0, u1, tau, diff_sampling, wh, d);
        w_hat[i] = wh;
        if (dw) dw[i] = d;
    }
}

// the device step counter of the in-kernel noise advances once per call (capturable: fresh noise per graph replay)
__global__ void sga_advance_kernel(unsigned long long* rng_step) { *rng_step += 1ull; }

}  // namespace shacira
