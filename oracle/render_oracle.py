"""TEST INFRASTRUCTURE ONLY -- float64 restatement of the packed ray integration the reference's tracer calls
(wisp/tracers/packed_rf_tracer.py:136-153 -> kaolin.render.spc.exponential_integration / sum_reduce, kaolin 0.13.0).

kaolin is absent from this image and from /root/reference (requirements: README.md:35), so this restates the
PUBLISHED algorithm -- per packed ray: T_i = exp(-exclusive cumsum of tau), w_i = T_i (1 - exp(-tau_i)),
ray_feats = sum_i w_i feats_i -- and is anchored on the reference's call site. PARITY WITH KAOLIN ITSELF IS UNPINNED.
Plain loops / torch float64; only tests/, smoke() and bench.py's cpu_baseline leg may import this."""
import numpy as np
import torch


def exponential_integration(feats, tau, boundary):
    """numpy float64: (ray_feats [R, NF], weights [S])."""
    feats = np.asarray(feats, dtype=np.float64)
    tau = np.asarray(tau, dtype=np.float64).reshape(-1)
    starts = list(np.nonzero(np.asarray(boundary).reshape(-1))[0]) + [tau.shape[0]]
    w = np.zeros_like(tau)
    out = np.zeros((len(starts) - 1, feats.shape[1]))
    for r in range(len(starts) - 1):
        acc = 0.0
        for i in range(starts[r], starts[r + 1]):
            w[i] = np.exp(-acc) * (1.0 - np.exp(-tau[i]))
            acc += tau[i]
            out[r] += w[i] * feats[i]
    return out, w


def exponential_integration_torch(feats, tau, boundary):
    """Differentiable float64 torch form of the same thing (autograd gives the reference gradients)."""
    tau = tau.reshape(-1)
    starts = torch.nonzero(boundary.reshape(-1)).squeeze(1).tolist() + [tau.shape[0]]
    outs, ws = [], []
    for r in range(len(starts) - 1):
        t = tau[starts[r]:starts[r + 1]]
        excl = torch.cumsum(t, 0) - t
        w = torch.exp(-excl) * (1.0 - torch.exp(-t))
        ws.append(w)
        outs.append((w.unsqueeze(1) * feats[starts[r]:starts[r + 1]]).sum(0))
    return torch.stack(outs), torch.cat(ws)


def voxel_samples(origins, dirs, ridx, depth, jitter):
    """float32 numpy restatement of OctreeAS._raymarch_voxel AFTER the kaolin raytrace
    (wisp/accelstructs/octree_as.py:195-228) with the reference's in-repo helpers
    (wisp/ops/spc/sampling.py:35-71) and the jitter injected:
        steps = (arange(K) + jitter) * (1 / K);  d = entry + (exit - entry) * steps           sampling.py:50-53
        deltas = d.diff(prepend=entry)                                                        octree_as.py:202
        samples = origins[ridx] + dirs[ridx] * d                                              octree_as.py:205-206
        boundary = expand_pack_boundary(mark_first_hit(ridx), K)                              octree_as.py:210-211
    Pinned against the reference's own sampling.py through tests/golden/sampling_ref.npz.
    Returns (ridx_out [M*K] int64, samples [M*K,3], depth_samples [M*K], deltas [M*K], boundary [M*K] bool)."""
    depth = np.asarray(depth, dtype=np.float32)
    jitter = np.asarray(jitter, dtype=np.float32)
    M, K = jitter.shape
    steps = np.arange(K, dtype=np.float32)[None].repeat(M, 0)
    steps = steps + jitter
    steps = steps * np.float32(1.0 / K)
    d = depth[:, 0:1] + (depth[:, 1:2] - depth[:, 0:1]) * steps
    deltas = np.diff(d, axis=1, prepend=depth[:, 0:1])
    ridx = np.asarray(ridx).astype(np.int64)
    o = np.asarray(origins, dtype=np.float32)[ridx][:, None]
    dr = np.asarray(dirs, dtype=np.float32)[ridx][:, None]
    samples = o + dr * d[..., None]
    first = np.ones(M, dtype=bool)
    first[1:] = ridx[1:] != ridx[:-1]
    boundary = np.zeros(M * K, dtype=bool)
    boundary[np.nonzero(first)[0] * K] = True
    return (np.repeat(ridx, K), samples.reshape(M * K, 3), d.reshape(-1), deltas.reshape(-1).astype(np.float32), boundary)


def raytrace_dense_bruteforce(occupancy, origins, dirs):
    """Ray / occupied-cell intersections by a slab test against EVERY occupied cell (no traversal logic at all), sorted
    by entry depth per ray: the independent yardstick for the 3D-DDA kernel that stands in for kaolin's
    unbatched_raytrace (OctreeAS.raytrace, wisp/accelstructs/octree_as.py:148-170; kaolin absent, parity unpinned).
    float64. Returns a list per ray of (cell index, t_enter, t_exit) with t_exit - t_enter > 0."""
    occ = np.asarray(occupancy).astype(bool)
    res = occ.shape[0]
    cs = 2.0 / res
    cells = np.argwhere(occ)
    lo = -1.0 + cells * cs
    hi = lo + cs
    out = []
    for o, d in zip(np.asarray(origins, dtype=np.float64), np.asarray(dirs, dtype=np.float64)):
        with np.errstate(divide="ignore", invalid="ignore"):
            t0 = (lo - o) / d
            t1 = (hi - o) / d
        par = d == 0.0
        inside = (o >= lo) & (o <= hi)
        tn = np.where(par, np.where(inside, -np.inf, np.inf), np.minimum(t0, t1)).max(1)
        tf = np.where(par, np.where(inside, np.inf, -np.inf), np.maximum(t0, t1)).min(1)
        tn = np.maximum(tn, 0.0)
        ok = tf > tn
        idx = (cells[:, 0] * res + cells[:, 1]) * res + cells[:, 2]
        hits = sorted(zip(tn[ok], tf[ok], idx[ok]))
        out.append([(int(c), float(a), float(b)) for a, b, c in hits])
    return out


def prune_dense(occupancy, density, jitter, density_decay, min_density):
    """float32 numpy restatement of NeuralRadianceField.prune on a dense grid (wisp/models/nefs/nerf.py:158-171), cell
    (x, y, z) at (x*res + y)*res + z, jitter and density injected:
        occupancy *= decay; samples = ((points + jitter) / res) * 2 - 1; occupancy = max(density, occupancy)
        mask = occupancy > min_density
    Returns (samples [res^3, 3], new occupancy [res^3], mask [res^3] bool)."""
    occ = np.asarray(occupancy, dtype=np.float32).reshape(-1)
    res = round(occ.shape[0] ** (1.0 / 3.0))
    g = np.arange(res)
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    samples = pts + np.asarray(jitter, dtype=np.float32)
    samples = samples / np.float32(res)
    samples = samples * np.float32(2.0) - np.float32(1.0)
    occ = occ * np.float32(density_decay)
    occ = np.maximum(np.asarray(density, dtype=np.float32).reshape(-1), occ)
    return samples, occ, occ > np.float32(min_density)
