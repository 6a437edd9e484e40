// render_kernels.cuh -- packed exponential integration along rays (the step AFTER the 3D grid
// and its decoders in the NeRF path).
//
// Reference call site: wisp/tracers/packed_rf_tracer.py:136-153
//     tau = density * deltas
//     ray_colors, transmittance = spc_render.exponential_integration(color, tau, boundary, exclusive=True)
//     alpha = spc_render.sum_reduce(transmittance, boundary)            (+ depth = sum_reduce(depths * transmittance))
// `spc_render` is kaolin.render.spc (kaolin==0.13.0, README.md:35): an absent third-party dependency. Its published
// algorithm, restated: samples are PACKED ray after ray (`boundary[i]` marks the first sample of a ray);
//     T_i = exp(-sum_{j<i, same ray} tau_j)          (exclusive cumsum)
//     w_i = T_i * (1 - exp(-tau_i))                  ("transmittance" returned per sample)
//     ray_feats_r = sum_{i in r} w_i * feats_i       (sum_reduce)
// Parity with kaolin itself is UNPINNED (package absent); the oracle (oracle/render_oracle.py) restates the same
// formulas in float64 and the gradients are checked against autograd of that restatement.
//
// One warp per ray: 32 samples per step, prefix sums by warp shuffles with a running carry. Forward: one pass.
// Backward: d w_k / d tau_i = -w_k for k > i and T_{i+1} for k = i, so
//     grad_tau_i = G_i * T_{i+1} - sum_{k>i} G_k w_k,   G_i = <grad_ray_feats_r, feats_i> + grad_w_i
// i.e. one forward sweep (T_{i+1}, parked in the output) and one reverse sweep (suffix sums).
#pragma once
#include "common.cuh"

namespace shacira {

constexpr int kRenderWarps = 8;

__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// ray_start[r] .. ray_start[r+1]: samples of ray r. feats [S, NF]; tau [S]; out: weights [S], ray_feats [R, NF],
// ray_alpha [R] (= sum of the weights; may be NULL).
template <int NF>
__global__ void __launch_bounds__(kRenderWarps * 32)
integrate_fwd_kernel(const float* __restrict__ feats, const float* __restrict__ tau, const int32_t* __restrict__ ray_start,
                     int32_t num_rays, float* __restrict__ weights, float* __restrict__ ray_feats,
                     float* __restrict__ ray_alpha) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * kRenderWarps + (threadIdx.x >> 5);
    if (r >= num_rays) return;
    const int s0 = ray_start[r], s1 = ray_start[r + 1];
    float carry = 0.0f, acc[NF], asum = 0.0f;
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = 0.0f;
    for (int base = s0; base < s1; base += 32) {
        const int i = base + lane;
        const float t = (i < s1) ? __ldg(tau + i) : 0.0f;
        const float incl = warp_incl_scan(t, lane);
        const float T = expf(-(carry + incl - t));
        const float w = T * (1.0f - expf(-t));
        if (i < s1) {
            weights[i] = w;
            asum += w;
#pragma unroll
            for (int f = 0; f < NF; ++f) acc[f] = fmaf(w, __ldg(feats + (int64_t)i * NF + f), acc[f]);
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const float v = warp_sum(acc[f]);
        if (lane == 0) ray_feats[(int64_t)r * NF + f] = v;
    }
    const float a = warp_sum(asum);
    if (lane == 0 && ray_alpha) ray_alpha[r] = a;
}

// grad_ray_feats [R, NF]; grad_weights [S] or NULL (gradient reaching the per-sample weights directly: alpha, depth);
// weights [S] from the forward. out: grad_feats [S, NF] (may be NULL), grad_tau [S].
template <int NF>
__global__ void __launch_bounds__(kRenderWarps * 32)
integrate_bwd_kernel(const float* __restrict__ feats, const float* __restrict__ tau, const float* __restrict__ weights,
                     const int32_t* __restrict__ ray_start, int32_t num_rays, const float* __restrict__ grad_ray_feats,
                     const float* __restrict__ grad_weights, float* __restrict__ grad_feats,
                     float* __restrict__ grad_tau) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * kRenderWarps + (threadIdx.x >> 5);
    if (r >= num_rays) return;
    const int s0 = ray_start[r], s1 = ray_start[r + 1];
    float gr[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) gr[f] = __ldg(grad_ray_feats + (int64_t)r * NF + f);
    // forward sweep: T_{i+1} = exp(-inclusive cumsum), parked in grad_tau
    float carry = 0.0f;
    for (int base = s0; base < s1; base += 32) {
        const int i = base + lane;
        const float t = (i < s1) ? __ldg(tau + i) : 0.0f;
        const float incl = warp_incl_scan(t, lane);
        if (i < s1) grad_tau[i] = expf(-(carry + incl));
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    // reverse sweep: suffix sums of G_k w_k
    float tail = 0.0f;  // sum over the chunks behind this one
    const int nchunks = (s1 - s0 + 31) >> 5;
    for (int c = nchunks - 1; c >= 0; --c) {
        const int i = s0 + 32 * c + lane;
        float G = 0.0f, w = 0.0f;
        if (i < s1) {
            w = weights[i];
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                G = fmaf(gr[f], __ldg(feats + (int64_t)i * NF + f), G);
                if (grad_feats) grad_feats[(int64_t)i * NF + f] = w * gr[f];
            }
            if (grad_weights) G += __ldg(grad_weights + i);
        }
        const float gw = G * w;
        const float incl = warp_incl_scan(gw, lane);
        const float total = __shfl_sync(0xffffffffu, incl, 31);
        const float after = tail + (total - incl);  // sum_{k > i} G_k w_k
        if (i < s1) grad_tau[i] = G * grad_tau[i] - after;
        tail += total;
    }
}

}  // namespace shacira

namespace shacira {

// Samples inside the intersected cells (the reference's "voxel" raymarch after the ray/cell intersection):
// OctreeAS._raymarch_voxel, wisp/accelstructs/octree_as.py:195-228, with sample_from_depth_intervals and
// expand_pack_boundary of wisp/ops/spc/sampling.py:35-71 -- ~10 PyTorch launches and four [M, K] temporaries there,
// one pass here. Nugget m = (ray ridx[m], entry/exit depth[m]); K samples each, jitter[m, k] in [0, 1) injected:
//   d = entry + (exit - entry) * ((k + jitter) * (1 / K))      separately rounded mul and add, as torch evaluates it
//   delta = d_k - d_{k-1} (d_{-1} = entry);  sample = origin + dir * d;  boundary = first sample of a ray's first nugget
__global__ void __launch_bounds__(256)
voxel_samples_kernel(const float* __restrict__ origins, const float* __restrict__ dirs, const int32_t* __restrict__ ridx,
                     const float* __restrict__ depth, const float* __restrict__ jitter, int64_t num_nuggets, int K,
                     float inv_k, int64_t* __restrict__ ridx_out, float* __restrict__ samples,
                     float* __restrict__ depth_samples, float* __restrict__ deltas, uint8_t* __restrict__ boundary) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= num_nuggets * K) return;
    const int64_t m = e / K;
    const int k = (int)(e - m * K);
    const float d0 = __ldg(depth + 2 * m), d1 = __ldg(depth + 2 * m + 1);
    const float span = __fsub_rn(d1, d0);
    const float st = __fmul_rn(__fadd_rn((float)k, __ldg(jitter + e)), inv_k);
    const float d = __fadd_rn(d0, __fmul_rn(span, st));
    float prev = d0;
    if (k > 0) prev = __fadd_rn(d0, __fmul_rn(span, __fmul_rn(__fadd_rn((float)(k - 1), __ldg(jitter + e - 1)), inv_k)));
    const int32_t r = __ldg(ridx + m);
    depth_samples[e] = d;
    deltas[e] = __fsub_rn(d, prev);
#pragma unroll
    for (int a = 0; a < 3; ++a)
        samples[e * 3 + a] = fmaf(__ldg(dirs + (int64_t)r * 3 + a), d, __ldg(origins + (int64_t)r * 3 + a));
    if (ridx_out) ridx_out[e] = r;
    boundary[e] = (k == 0 && (m == 0 || __ldg(ridx + m - 1) != r)) ? 1 : 0;
}

}  // namespace shacira

namespace shacira {

// Ray / occupied-cell intersections against a DENSE occupancy grid (res^3 cells over [-1, 1]^3, cell (x, y, z) at
// (x * res + y) * res + z): stands in for kaolin's `spc_render.unbatched_raytrace(octree, ..., level,
// return_depth=True, with_exit=True)` as called by OctreeAS.raytrace (wisp/accelstructs/octree_as.py:148-170) when the
// structure is the dense grid the reference builds for its hash-grid NeRFs (OctreeAS.make_dense, :119-127) or a
// pruned copy of it (nerf.py:150-185). kaolin is absent: this is a from-scratch 3D-DDA (Amanatides & Woo); parity with
// kaolin's traversal order on grazing rays is unpinned, the oracle is a brute-force slab test per occupied cell.
// One thread per ray, two passes (count -> exclusive scan on the host side of the ABI -> fill): nuggets come out
// packed ray after ray, sorted by depth along each ray, which is what the rest of the path expects.
struct DdaState {
    int cell[3], step[3];
    float tmax[3], tdelta[3];
    float t, tfar;
    bool hit;
};

__device__ __forceinline__ void dda_init(const float* o, const float* d, int res, DdaState& s) {
    float tn = 0.0f, tf = 3.0e38f;
    float inv[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        inv[a] = d[a] != 0.0f ? 1.0f / d[a] : 3.0e38f;
        if (d[a] != 0.0f) {
            const float t0 = (-1.0f - o[a]) * inv[a], t1 = (1.0f - o[a]) * inv[a];
            tn = fmaxf(tn, fminf(t0, t1));
            tf = fminf(tf, fmaxf(t0, t1));
        } else if (o[a] < -1.0f || o[a] > 1.0f) {
            tf = -1.0f;  // parallel to the slab and outside it
        }
    }
    s.hit = tn < tf;
    s.t = tn;
    s.tfar = tf;
    if (!s.hit) return;
    const float cs = 2.0f / (float)res;
    const float te = tn + 1e-6f * fmaxf(1.0f, tn);  // just inside the box
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float p = fmaf(d[a], te, o[a]);
        int c = (int)floorf((p + 1.0f) * 0.5f * (float)res);
        c = min(max(c, 0), res - 1);
        s.cell[a] = c;
        s.step[a] = d[a] > 0.0f ? 1 : (d[a] < 0.0f ? -1 : 0);
        if (d[a] != 0.0f) {
            const float bound = -1.0f + (float)(c + (d[a] > 0.0f ? 1 : 0)) * cs;
            s.tmax[a] = (bound - o[a]) * inv[a];
            s.tdelta[a] = cs * fabsf(inv[a]);
        } else {
            s.tmax[a] = 3.0e38f;
            s.tdelta[a] = 3.0e38f;
        }
    }
}

// FILL = false: count[r] = number of occupied cells ray r crosses. FILL = true: write them at offset[r].
template <bool FILL>
__global__ void __launch_bounds__(128)
raytrace_dense_kernel(const uint8_t* __restrict__ occ, int res, const float* __restrict__ origins,
                      const float* __restrict__ dirs, int32_t num_rays, int32_t* __restrict__ count,
                      const int64_t* __restrict__ offset, int32_t* __restrict__ ridx, int32_t* __restrict__ pidx,
                      float* __restrict__ depth) {
    const int r = blockIdx.x * 128 + threadIdx.x;
    if (r >= num_rays) return;
    const float o[3] = {origins[r * 3], origins[r * 3 + 1], origins[r * 3 + 2]};
    const float d[3] = {dirs[r * 3], dirs[r * 3 + 1], dirs[r * 3 + 2]};
    DdaState s;
    dda_init(o, d, res, s);
    int n = 0;
    int64_t w = FILL ? offset[r] : 0;
    if (s.hit) {
        for (int guard = 0; guard < 3 * res + 3; ++guard) {
            const int ax = (s.tmax[0] <= s.tmax[1]) ? (s.tmax[0] <= s.tmax[2] ? 0 : 2) : (s.tmax[1] <= s.tmax[2] ? 1 : 2);
            const float texit = fminf(s.tmax[ax], s.tfar);
            const int cidx = (s.cell[0] * res + s.cell[1]) * res + s.cell[2];
            if (texit > s.t && occ[cidx]) {
                if (FILL) {
                    ridx[w] = r;
                    pidx[w] = cidx;
                    depth[2 * w] = s.t;
                    depth[2 * w + 1] = texit;
                    ++w;
                }
                ++n;
            }
            if (s.tmax[ax] >= s.tfar) break;
            s.t = texit;
            s.cell[ax] += s.step[ax];
            if (s.cell[ax] < 0 || s.cell[ax] >= res) break;
            s.tmax[ax] += s.tdelta[ax];
        }
    }
    if (!FILL) count[r] = n;
}

}  // namespace shacira

namespace shacira {

// Occupancy pruning on the dense grid (NeuralRadianceField.prune, wisp/models/nefs/nerf.py:150-185): one jittered
// sample per cell, then occupancy = max(density, occupancy * decay) and the cells above `min_density` stay.
// Cell e = (x * res + y) * res + z, the layout raytrace_dense_kernel reads.
//   sample = ((cell + jitter) / res) * 2 - 1                       nerf.py:160-165
__global__ void __launch_bounds__(256)
prune_samples_kernel(int res, const float* __restrict__ jitter, float* __restrict__ samples) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t cells = (int64_t)res * res * res;
    if (e >= cells) return;
    const int c[3] = {(int)(e / ((int64_t)res * res)), (int)((e / res) % res), (int)(e % res)};
    const float r = (float)res;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float s = __fdiv_rn(__fadd_rn((float)c[a], __ldg(jitter + e * 3 + a)), r);
        samples[e * 3 + a] = __fsub_rn(__fmul_rn(s, 2.0f), 1.0f);
    }
}
//   occupancy = max(density, occupancy * decay);  mask = occupancy > min_density        nerf.py:158,169-171
__global__ void __launch_bounds__(256)
prune_update_kernel(int64_t cells, const float* __restrict__ density, float decay, float min_density,
                    float* __restrict__ occupancy, uint8_t* __restrict__ mask) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= cells) return;
    const float o = fmaxf(__ldg(density + e), __fmul_rn(occupancy[e], decay));
    occupancy[e] = o;
    mask[e] = o > min_density ? 1 : 0;
}

}  // namespace shacira
