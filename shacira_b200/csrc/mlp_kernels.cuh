// mlp_kernels.cuh -- SURVEY section 8 row f-1: the tiny decoder MLP that consumes the grid features, fused with
// the image loss: forward, MSE, backward to the features AND all weight gradients in ONE pass over the points.
//
// Reference: NeuralImage.rgb -> BasicDecoder (wisp/models/nefs/image.py:109-120,152; hidden 16, two hidden layers,
// ReLU, bias; wisp/models/decoders/basic_decoders.py:60-100) followed by ((pred - gt)**2).mean()
// (wisp/trainers/image_trainer.py:298-300). In PyTorch that is ~15 kernels per step, each streaming an
// [N, 16] activation through HBM/L2; here the activations never leave registers / shared memory:
// per point 64 B features in, 12 B target in, 64 B feature gradient out (+12 B prediction when asked for).
//
//   y  = W3 relu(W2 relu(W1 x + b1) + b2) + b3            (torch.nn.Linear layout: W[out][in])
//   L  = sum((y - gt)^2) / (N * OUT)
//
// One lane owns one point for forward + backward (weights broadcast from shared memory as 16-byte vectors);
// the weight gradients dW = sum_p a_p b_p^T are formed per warp from the 32 points' activations staged in shared
// memory, each lane accumulating an 8-wide slice in registers across ALL its warp's points (persistent CTAs),
// then reduced over the block in shared memory and added to global memory once per CTA and value.
#pragma once
#include "common.cuh"

namespace shacira {

constexpr int kMlpThreads = 256;
constexpr int kMlpWarps = kMlpThreads / 32;

template <int IN, int H, int OUT>
struct MlpSmem {
    // weights, and their transposes for the backward products (row = reduction index -> 16-byte broadcast loads)
    float W1[H][IN], W2[H][H], W3[OUT][H];
    float W1t[IN][H], W2t[H][H], W3t[H][4];  // W3t padded to 4 outputs
    float b1[H], b2[H], b3[4];
    // per-warp staging of the 32 points' activations for the weight-gradient products
    float x[kMlpWarps][32][IN + 4];
    float h1[kMlpWarps][32][H + 4], h2[kMlpWarps][32][H + 4];
    float d1[kMlpWarps][32][H + 4], d2[kMlpWarps][32][H + 4];
    float dy[kMlpWarps][32][4];
    // block reduction of the per-warp gradient slices
    float gW1[H * IN], gW2[H * H], gW3[OUT * H], gb1[H], gb2[H], gb3[4];
    double loss;
};

template <int IN, int H, int OUT>
__global__ void __launch_bounds__(kMlpThreads, 2)
mlp_mse_step_kernel(const float* __restrict__ x, const float* __restrict__ gt, int64_t n, const float* __restrict__ W1,
                    const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                    const float* __restrict__ W3, const float* __restrict__ b3, float grad_scale /* 2 / (N*OUT) */,
                    float* __restrict__ gx, float* __restrict__ pred, double* __restrict__ loss_sum,
                    float* __restrict__ grad_params) {
    // packed gradient buffer: W1 [H][IN] | b1 [H] | W2 [H][H] | b2 [H] | W3 [OUT][H] | b3 [OUT]
    float* gW1 = grad_params;
    float* gb1 = gW1 + H * IN;
    float* gW2 = gb1 + H;
    float* gb2 = gW2 + H * H;
    float* gW3 = gb2 + H;
    float* gb3 = gW3 + OUT * H;
    static_assert(IN % 8 == 0 && H == 16 && OUT <= 4, "shape");
    extern __shared__ __align__(16) unsigned char s_raw[];
    MlpSmem<IN, H, OUT>& S = *reinterpret_cast<MlpSmem<IN, H, OUT>*>(s_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < H * IN; e += kMlpThreads) {
        const float w = W1[e];
        S.W1[e / IN][e % IN] = w;
        S.W1t[e % IN][e / IN] = w;
        S.gW1[e] = 0.0f;
    }
    for (int e = tid; e < H * H; e += kMlpThreads) {
        const float w = W2[e];
        S.W2[e / H][e % H] = w;
        S.W2t[e % H][e / H] = w;
        S.gW2[e] = 0.0f;
    }
    for (int e = tid; e < H * 4; e += kMlpThreads) S.W3t[e / 4][e % 4] = 0.0f;
    __syncthreads();
    for (int e = tid; e < OUT * H; e += kMlpThreads) {
        const float w = W3[e];
        S.W3[e / H][e % H] = w;
        S.W3t[e % H][e / H] = w;
        S.gW3[e] = 0.0f;
    }
    if (tid < H) { S.b1[tid] = b1[tid]; S.b2[tid] = b2[tid]; S.gb1[tid] = 0.0f; S.gb2[tid] = 0.0f; }
    if (tid < 4) { S.b3[tid] = tid < OUT ? b3[tid] : 0.0f; S.gb3[tid] = 0.0f; }
    if (tid == 0) S.loss = 0.0;
    __syncthreads();

    // register slices of the weight gradients owned by this lane (summed over all points of this warp)
    //   dW1[i][m0..m0+IN/2)  i = lane/2, m0 = (lane&1) * IN/2         (H*IN/32 values per lane)
    //   dW2[i][j0..j0+8)     i = lane/2, j0 = (lane&1) * 8
    //   dW3[k][j]            lanes < OUT*H/2 own two values; biases: lane < H owns db1[lane], db2[lane]; lane < OUT db3
    constexpr int S1 = IN / 2;
    float aW1[S1], aW2[8], aW3[2] = {0.0f, 0.0f}, ab1 = 0.0f, ab2 = 0.0f, ab3 = 0.0f;
#pragma unroll
    for (int e = 0; e < S1; ++e) aW1[e] = 0.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) aW2[e] = 0.0f;
    float my_loss = 0.0f;
    const int wi = lane >> 1, wj0 = (lane & 1) * 8, wm0 = (lane & 1) * S1;

    const int64_t warps_total = (int64_t)gridDim.x * kMlpWarps;
    for (int64_t base = ((int64_t)blockIdx.x * kMlpWarps + warp) * 32; base < n; base += warps_total * 32) {
        const int64_t p = base + lane;
        const bool live = p < n;
        float xi[IN], h1[H], h2[H], y[4] = {0.0f, 0.0f, 0.0f, 0.0f}, tgt[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int e = 0; e < IN; ++e) xi[e] = 0.0f;
        if (live) {
#pragma unroll
            for (int q = 0; q < IN / 4; ++q) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * IN) + q);
                xi[4 * q] = v.x; xi[4 * q + 1] = v.y; xi[4 * q + 2] = v.z; xi[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int k = 0; k < OUT; ++k) tgt[k] = __ldg(gt + p * OUT + k);
        }
        // ---- forward ----
#pragma unroll
        for (int i = 0; i < H; ++i) {
            float acc = S.b1[i];
#pragma unroll
            for (int q = 0; q < IN / 4; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(&S.W1[i][4 * q]);
                acc = fmaf(xi[4 * q], w.x, acc); acc = fmaf(xi[4 * q + 1], w.y, acc);
                acc = fmaf(xi[4 * q + 2], w.z, acc); acc = fmaf(xi[4 * q + 3], w.w, acc);
            }
            h1[i] = fmaxf(acc, 0.0f);
        }
#pragma unroll
        for (int j = 0; j < H; ++j) {
            float acc = S.b2[j];
#pragma unroll
            for (int q = 0; q < H / 4; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(&S.W2[j][4 * q]);
                acc = fmaf(h1[4 * q], w.x, acc); acc = fmaf(h1[4 * q + 1], w.y, acc);
                acc = fmaf(h1[4 * q + 2], w.z, acc); acc = fmaf(h1[4 * q + 3], w.w, acc);
            }
            h2[j] = fmaxf(acc, 0.0f);
        }
#pragma unroll
        for (int k = 0; k < OUT; ++k) {
            float acc = S.b3[k];
#pragma unroll
            for (int q = 0; q < H / 4; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(&S.W3[k][4 * q]);
                acc = fmaf(h2[4 * q], w.x, acc); acc = fmaf(h2[4 * q + 1], w.y, acc);
                acc = fmaf(h2[4 * q + 2], w.z, acc); acc = fmaf(h2[4 * q + 3], w.w, acc);
            }
            y[k] = acc;
        }
        // ---- loss and its gradient ----
        float dy[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (live) {
#pragma unroll
            for (int k = 0; k < OUT; ++k) {
                const float e = y[k] - tgt[k];
                my_loss = fmaf(e, e, my_loss);
                dy[k] = e * grad_scale;
                if (pred) pred[p * OUT + k] = y[k];
            }
        }
        // ---- backward to the features ----
        float d2[H], d1[H];
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const float4 w = *reinterpret_cast<const float4*>(&S.W3t[j][0]);
            float acc = dy[0] * w.x;
            acc = fmaf(dy[1], w.y, acc); acc = fmaf(dy[2], w.z, acc); acc = fmaf(dy[3], w.w, acc);
            d2[j] = h2[j] > 0.0f ? acc : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < H; ++i) {
            float acc = 0.0f;
#pragma unroll
            for (int q = 0; q < H / 4; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(&S.W2t[i][4 * q]);  // W2[j][i], j = 4q..
                acc = fmaf(d2[4 * q], w.x, acc); acc = fmaf(d2[4 * q + 1], w.y, acc);
                acc = fmaf(d2[4 * q + 2], w.z, acc); acc = fmaf(d2[4 * q + 3], w.w, acc);
            }
            d1[i] = h1[i] > 0.0f ? acc : 0.0f;
        }
        if (live) {
#pragma unroll
            for (int q = 0; q < IN / 4; ++q) {
                float o[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int m = 4 * q + r;
                    float acc = 0.0f;
#pragma unroll
                    for (int qq = 0; qq < H / 4; ++qq) {
                        const float4 w = *reinterpret_cast<const float4*>(&S.W1t[m][4 * qq]);  // W1[i][m], i = 4qq..
                        acc = fmaf(d1[4 * qq], w.x, acc); acc = fmaf(d1[4 * qq + 1], w.y, acc);
                        acc = fmaf(d1[4 * qq + 2], w.z, acc); acc = fmaf(d1[4 * qq + 3], w.w, acc);
                    }
                    o[r] = acc;
                }
                reinterpret_cast<float4*>(gx + p * IN)[q] = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        // ---- weight gradients: stage this warp's 32 points, every lane accumulates its slice ----
        __syncwarp();
        // 16-byte stores, row stride (IN|H)+4 floats = 20 banks: a quarter-warp covers all 32 banks, no conflicts
#pragma unroll
        for (int q = 0; q < IN / 4; ++q)
            *reinterpret_cast<float4*>(&S.x[warp][lane][4 * q]) = make_float4(xi[4 * q], xi[4 * q + 1], xi[4 * q + 2], xi[4 * q + 3]);
#pragma unroll
        for (int q = 0; q < H / 4; ++q) {
            *reinterpret_cast<float4*>(&S.h1[warp][lane][4 * q]) = make_float4(h1[4 * q], h1[4 * q + 1], h1[4 * q + 2], h1[4 * q + 3]);
            *reinterpret_cast<float4*>(&S.h2[warp][lane][4 * q]) = make_float4(h2[4 * q], h2[4 * q + 1], h2[4 * q + 2], h2[4 * q + 3]);
            *reinterpret_cast<float4*>(&S.d1[warp][lane][4 * q]) = make_float4(d1[4 * q], d1[4 * q + 1], d1[4 * q + 2], d1[4 * q + 3]);
            *reinterpret_cast<float4*>(&S.d2[warp][lane][4 * q]) = make_float4(d2[4 * q], d2[4 * q + 1], d2[4 * q + 2], d2[4 * q + 3]);
        }
        *reinterpret_cast<float4*>(&S.dy[warp][lane][0]) = make_float4(dy[0], dy[1], dy[2], dy[3]);
        __syncwarp();
#pragma unroll 4
        for (int q = 0; q < 32; ++q) {
            const float d1q = S.d1[warp][q][wi];   // dW1[i][m] += d1[i] * x[m]
#pragma unroll
            for (int e = 0; e < S1; e += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&S.x[warp][q][wm0 + e]);
                aW1[e] = fmaf(d1q, v.x, aW1[e]); aW1[e + 1] = fmaf(d1q, v.y, aW1[e + 1]);
                aW1[e + 2] = fmaf(d1q, v.z, aW1[e + 2]); aW1[e + 3] = fmaf(d1q, v.w, aW1[e + 3]);
            }
            const float d2q = S.d2[warp][q][wi];   // dW2[j][i'] += d2[j] * h1[i'] : row j = wi, columns wj0..wj0+8
#pragma unroll
            for (int e = 0; e < 8; e += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&S.h1[warp][q][wj0 + e]);
                aW2[e] = fmaf(d2q, v.x, aW2[e]); aW2[e + 1] = fmaf(d2q, v.y, aW2[e + 1]);
                aW2[e + 2] = fmaf(d2q, v.z, aW2[e + 2]); aW2[e + 3] = fmaf(d2q, v.w, aW2[e + 3]);
            }
            if (lane < OUT * H / 2) {               // dW3[k][j] += dy[k] * h2[j] : two values per lane
                const int v0 = 2 * lane, k0 = v0 / H, j0 = v0 % H;
                const float dyk = S.dy[warp][q][k0];
                aW3[0] = fmaf(dyk, S.h2[warp][q][j0], aW3[0]);
                aW3[1] = fmaf(dyk, S.h2[warp][q][j0 + 1], aW3[1]);
            }
            if (lane < H) {
                ab1 += S.d1[warp][q][lane];
                ab2 += S.d2[warp][q][lane];
            }
            if (lane < OUT) ab3 += S.dy[warp][q][lane];
        }
    }
    // ---- block reduction (shared memory, 8 warps) then one global add per CTA and value ----
#pragma unroll
    for (int e = 0; e < S1; ++e) atomicAdd(&S.gW1[wi * IN + wm0 + e], aW1[e]);
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&S.gW2[wi * H + wj0 + e], aW2[e]);
    if (lane < OUT * H / 2) {
        atomicAdd(&S.gW3[2 * lane], aW3[0]);
        atomicAdd(&S.gW3[2 * lane + 1], aW3[1]);
    }
    if (lane < H) { atomicAdd(&S.gb1[lane], ab1); atomicAdd(&S.gb2[lane], ab2); }
    if (lane < OUT) atomicAdd(&S.gb3[lane], ab3);
    const float wl = warp_sum(my_loss);
    if (lane == 0) atomicAdd(&S.loss, (double)wl);
    __syncthreads();
    for (int e = tid; e < H * IN; e += kMlpThreads) red_add(gW1 + e, S.gW1[e]);
    for (int e = tid; e < H * H; e += kMlpThreads) red_add(gW2 + e, S.gW2[e]);
    for (int e = tid; e < OUT * H; e += kMlpThreads) red_add(gW3 + e, S.gW3[e]);
    if (tid < H) { red_add(gb1 + tid, S.gb1[tid]); red_add(gb2 + tid, S.gb2[tid]); }
    if (tid < OUT) red_add(gb3 + tid, S.gb3[tid]);
    if (tid == 0) atomicAdd(loss_sum, S.loss);
}

}  // namespace shacira
