// optimizer_kernels.cuh -- the whole optimizer side of the image-fit step as ONE launch.
//
// Reference: optimizer.step() over the parameter groups of base_trainer.py:206-266 (image_trainer.py:321-359) and, at the
// start of the next step, the SGA sample of the latents (basic_latent_decoder.py:183-191). Here:
//   CTAs [0, S.num)      Adam over the ~20 small tensors (decoder MLP, latent decoder, density model) with their chain
//                        rules, exactly multi_adam_kernel's per-CTA work;
//   the remaining CTAs   Adam over the latent table (gradient = grid gradient x d w_hat / d w + (lambda / rows) x bit-rate
//                        gradient, adam_step_sum_kernel's math) and, with SGA on, the NEXT step's sample of the freshly
//                        updated latents (w_hat, d w_hat / d w) while they are still in registers -- the table is read and
//                        written once per step instead of three times (SGA kernel, Adam kernel) and two launches go away;
//   the last CTA         (arrival ticket) advances the step counters and the SGA draw counter.
#pragma once
#include "entropy_kernels.cuh"
#include "sga_kernels.cuh"

namespace shacira {

struct TableAdam {
    float* p;                 // latents [n]
    float* g;                 // grid gradient (cleared after use)
    const float* gmul;        // d w_hat / d w of THIS step's sample (SGA) or NULL
    const float* g2;          // bit-rate gradient or NULL
    const float* scale2;      // device scalar lambda or NULL
    float mul2;               // 1 / rows
    float* m;
    float* v;
    int64_t n;
    float lr, weight_decay;
    // next step's SGA sample (w_hat == NULL: off)
    const float* temperature;
    int diff_sampling;
    unsigned long long seed;
    unsigned long long* rng_step;   // index of the next draw; advanced by the last CTA
    float* w_hat;
    float* dw;
};

__global__ void __launch_bounds__(256)
fit_optimizer_kernel(const __grid_constant__ AdamSegs S, const __grid_constant__ TableAdam Tb, float beta1, float beta2,
                     float eps, float* __restrict__ step_small, float* __restrict__ step_table,
                     const float* __restrict__ scale, const float* __restrict__ div, float* __restrict__ A_out, int C, int F,
                     unsigned* __restrict__ ticket) {
    if ((int)blockIdx.x < S.num) {
        const float t = *step_small + 1.0f;
        const float bc1 = 1.0f - powf(beta1, t), inv_sqrt_bc2 = rsqrtf(1.0f - powf(beta2, t));
        const shacira_adam_seg_t& sg = S.seg[blockIdx.x];
        const float gs = sg.grad_mul * (sg.grad_scale ? *sg.grad_scale : 1.0f);
        const float step_size = sg.lr / bc1;
        for (int i = threadIdx.x; i < sg.n; i += 256) {
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
        __syncthreads();
        if (A_out && sg.param == scale)
            for (int e = threadIdx.x; e < C * F; e += 256) A_out[e] = scale[e] / div[e / F];
    } else {
        const float t = *step_table + 1.0f;
        const float bc1 = 1.0f - powf(beta1, t), inv_sqrt_bc2 = rsqrtf(1.0f - powf(beta2, t));
        const float step_size = Tb.lr / bc1;
        const float s2 = Tb.g2 ? (Tb.scale2 ? *Tb.scale2 * Tb.mul2 : Tb.mul2) : 0.0f;
        const bool sga = Tb.w_hat != nullptr;
        float tau = 1.0f;
        uint32_t base = 0u;
        if (sga) {
            tau = __ldg(Tb.temperature);
            base = sga_rng_base(Tb.rng_step ? *Tb.rng_step : 0ull, Tb.seed);
        }
        const int64_t i4 = ((int64_t)(blockIdx.x - S.num) * 256 + threadIdx.x) * 4;
        const int cnt = (int)max((int64_t)0, min((int64_t)4, Tb.n - i4));
        const bool vec = cnt == 4 && ((reinterpret_cast<uintptr_t>(Tb.p) | reinterpret_cast<uintptr_t>(Tb.g) |
                                       reinterpret_cast<uintptr_t>(Tb.g2) | reinterpret_cast<uintptr_t>(Tb.m) |
                                       reinterpret_cast<uintptr_t>(Tb.v) | reinterpret_cast<uintptr_t>(Tb.gmul) |
                                       reinterpret_cast<uintptr_t>(Tb.w_hat) | reinterpret_cast<uintptr_t>(Tb.dw)) & 15) == 0;
        float P[4], M[4], V[4], G[4], Q[4], H[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { P[k] = M[k] = V[k] = G[k] = H[k] = 0.0f; Q[k] = 1.0f; }
        if (vec) {
            *reinterpret_cast<float4*>(P) = *reinterpret_cast<const float4*>(Tb.p + i4);
            *reinterpret_cast<float4*>(M) = *reinterpret_cast<const float4*>(Tb.m + i4);
            *reinterpret_cast<float4*>(V) = *reinterpret_cast<const float4*>(Tb.v + i4);
            *reinterpret_cast<float4*>(G) = *reinterpret_cast<const float4*>(Tb.g + i4);
            if (Tb.gmul) *reinterpret_cast<float4*>(Q) = *reinterpret_cast<const float4*>(Tb.gmul + i4);
            if (Tb.g2) *reinterpret_cast<float4*>(H) = *reinterpret_cast<const float4*>(Tb.g2 + i4);
        } else {
            for (int k = 0; k < cnt; ++k) {
                P[k] = Tb.p[i4 + k]; M[k] = Tb.m[i4 + k]; V[k] = Tb.v[i4 + k]; G[k] = Tb.g[i4 + k];
                if (Tb.gmul) Q[k] = Tb.gmul[i4 + k];
                if (Tb.g2) H[k] = Tb.g2[i4 + k];
            }
        }
        float WH[4], DW[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gi = Tb.gmul ? G[k] * Q[k] : G[k];
            if (Tb.g2) gi = fmaf(s2, H[k], gi);
            const float gk = fmaf(Tb.weight_decay, P[k], gi);
            M[k] = fmaf(beta1, M[k], (1.0f - beta1) * gk);
            V[k] = fmaf(beta2, V[k], (1.0f - beta2) * gk * gk);
            P[k] -= step_size * M[k] / (sqrtf(V[k]) * inv_sqrt_bc2 + eps);
            WH[k] = DW[k] = 0.0f;
            if (sga && k < cnt) {
                const int64_t i = i4 + k;
                sga_sample(P[k], sga_uniform((uint32_t)(2 * i), base), sga_uniform((uint32_t)(2 * i + 1), base), tau,
                           Tb.diff_sampling, WH[k], DW[k]);
            }
        }
        if (vec) {
            *reinterpret_cast<float4*>(Tb.p + i4) = *reinterpret_cast<float4*>(P);
            *reinterpret_cast<float4*>(Tb.m + i4) = *reinterpret_cast<float4*>(M);
            *reinterpret_cast<float4*>(Tb.v + i4) = *reinterpret_cast<float4*>(V);
            *reinterpret_cast<float4*>(Tb.g + i4) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (sga) {
                *reinterpret_cast<float4*>(Tb.w_hat + i4) = *reinterpret_cast<float4*>(WH);
                if (Tb.dw) *reinterpret_cast<float4*>(Tb.dw + i4) = *reinterpret_cast<float4*>(DW);
            }
        } else {
            for (int k = 0; k < cnt; ++k) {
                Tb.p[i4 + k] = P[k]; Tb.m[i4 + k] = M[k]; Tb.v[i4 + k] = V[k]; Tb.g[i4 + k] = 0.0f;
                if (sga) { Tb.w_hat[i4 + k] = WH[k]; if (Tb.dw) Tb.dw[i4 + k] = DW[k]; }
            }
        }
    }
    // every CTA has read the counters when the last one arrives: it advances them
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *step_small += 1.0f;
            *step_table += 1.0f;
            if (Tb.w_hat && Tb.rng_step) *Tb.rng_step += 1ull;
            *ticket = 0u;
        }
    }
}

}  // namespace shacira
