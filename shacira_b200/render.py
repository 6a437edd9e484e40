"""Packed ray integration: the `kaolin.render.spc` calls of the reference's tracer, on the B200 kernels
(SURVEY section 8 row f-3).

`wisp/tracers/packed_rf_tracer.py:136-153` does, right after the 3D grid and its decoders:

    tau = density * deltas
    ray_colors, transmittance = spc_render.exponential_integration(color, tau, boundary, exclusive=True)
    alpha = spc_render.sum_reduce(transmittance, boundary)
    ray_depth = spc_render.sum_reduce(depths * transmittance, boundary)

Same names and argument meaning here (`boundary`: bool [S], True at the first sample of every packed ray). kaolin
(0.13.0) is an absent dependency: its published algorithm is restated, parity with the package itself is unpinned
(oracle/render_oracle.py is the float64 restatement the tests check against). No CPU fallback.
"""
import torch

from . import _lib


def ray_starts(boundary):
    """int32 [R + 1]: first sample of every ray, then the sample count (kaolin's pack boundaries as offsets)."""
    if boundary.dtype != torch.bool:
        boundary = boundary != 0
    S = boundary.shape[0]
    starts = torch.nonzero(boundary, as_tuple=False).squeeze(1).to(torch.int32)
    return torch.cat((starts, torch.tensor([S], dtype=torch.int32, device=boundary.device)))


class _ExponentialIntegration(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, tau, starts, tau_2d):
        lib = _lib.load()
        feats, tau = _lib._f32c(feats, "feats"), _lib._f32c(tau.reshape(-1), "tau")
        S, NF = feats.shape
        R = starts.numel() - 1
        weights = torch.empty((S,), dtype=torch.float32, device=feats.device)
        ray_feats = torch.empty((R, NF), dtype=torch.float32, device=feats.device)
        with torch.cuda.device(feats.device):
            _lib._check(lib.shacira_integrate_forward(_lib._ptr(feats), _lib._ptr(tau), _lib._ptr(starts), R, NF,
                                                      _lib._ptr(weights), _lib._ptr(ray_feats), None, _lib._stream()))
        ctx.save_for_backward(feats, tau, weights, starts)
        ctx.tau_2d = tau_2d
        return ray_feats, weights.unsqueeze(1)

    @staticmethod
    def backward(ctx, g_ray, g_w):
        feats, tau, weights, starts = ctx.saved_tensors
        lib = _lib.load()
        S, NF = feats.shape
        R = starts.numel() - 1
        g_ray = _lib._f32c(g_ray, "grad_ray_feats")
        g_w = _lib._f32c(g_w.reshape(-1), "grad_weights") if g_w is not None else None
        g_feats = torch.empty_like(feats) if ctx.needs_input_grad[0] else None
        g_tau = torch.empty_like(tau)
        with torch.cuda.device(feats.device):
            _lib._check(lib.shacira_integrate_backward(_lib._ptr(feats), _lib._ptr(tau), _lib._ptr(weights),
                                                       _lib._ptr(starts), R, NF, _lib._ptr(g_ray), _lib._ptr(g_w),
                                                       _lib._ptr(g_feats), _lib._ptr(g_tau), _lib._stream()))
        return g_feats, (g_tau.unsqueeze(1) if ctx.tau_2d else g_tau), None, None


def exponential_integration(feats, tau, boundary, exclusive=True, starts=None):
    """(ray_feats [R, NF], transmittance weights [S, 1]) -- spc_render.exponential_integration. Pass `starts`
    (ray_starts(boundary)) to reuse the offsets across calls; exclusive=False is not used by the reference."""
    if not exclusive:
        raise _lib.ShaciraError(_lib.ERR_UNSUPPORTED, "exponential_integration: exclusive=False is not implemented")
    if starts is None:
        starts = ray_starts(boundary)
    return _ExponentialIntegration.apply(feats, tau, starts, tau.dim() == 2)


def sum_reduce(feats, boundary, starts=None):
    """Per-ray sums of packed per-sample values [S, F] -> [R, F] (spc_render.sum_reduce)."""
    if starts is None:
        starts = ray_starts(boundary)
    seg = torch.zeros(feats.shape[0], dtype=torch.int64, device=feats.device)
    seg[starts[1:-1].long()] = 1
    seg = torch.cumsum(seg, 0)
    out = torch.zeros((starts.numel() - 1, feats.shape[1]), dtype=feats.dtype, device=feats.device)
    return out.index_add_(0, seg, feats)


def voxel_samples(origins, dirs, ridx, depth, num_samples, jitter=None):
    """Everything `OctreeAS._raymarch_voxel` does after the ray/cell intersection (octree_as.py:195-228), one kernel:
    (ridx [M*K] int64, samples [M*K, 3], depth_samples [M*K, 1], deltas [M*K, 1], boundary [M*K] bool) from the
    intersection "nuggets" (ridx [M] sorted by ray, depth [M, 2] = entry/exit). `jitter` [M, K] in [0, 1) is the
    stratified-sampling draw (the reference's `torch.rand_like`, sampling.py:51); drawn on the device when None."""
    lib = _lib.load()
    origins, dirs, depth = _lib._f32c(origins, "origins"), _lib._f32c(dirs, "dirs"), _lib._f32c(depth, "depth")
    if not ridx.is_cuda:
        raise _lib.ShaciraError(_lib.ERR_INVALID_ARGUMENT, "ridx must be a CUDA tensor (no CPU fallback)")
    ridx32 = ridx.to(torch.int32).contiguous()
    M, K = depth.shape[0], int(num_samples)
    dev = depth.device
    if jitter is None:
        jitter = torch.rand((M, K), dtype=torch.float32, device=dev)
    jitter = _lib._f32c(jitter, "jitter")
    ridx_out = torch.empty((M * K,), dtype=torch.int64, device=dev)
    samples = torch.empty((M * K, 3), dtype=torch.float32, device=dev)
    depth_samples = torch.empty((M * K, 1), dtype=torch.float32, device=dev)
    deltas = torch.empty((M * K, 1), dtype=torch.float32, device=dev)
    boundary = torch.empty((M * K,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib._check(lib.shacira_voxel_samples(_lib._ptr(origins), _lib._ptr(dirs), _lib._ptr(ridx32), _lib._ptr(depth),
                                              _lib._ptr(jitter), M, K, _lib._ptr(ridx_out), _lib._ptr(samples),
                                              _lib._ptr(depth_samples), _lib._ptr(deltas), _lib._ptr(boundary),
                                              _lib._stream()))
    return ridx_out, samples, depth_samples, deltas, boundary.bool()


def raytrace_dense(occupancy, origins, dirs):
    """(ridx int32 [M], pidx int32 [M], depth [M, 2]) -- the fields of `OctreeAS.raytrace(rays, level, with_exit=True)`
    (octree_as.py:148-170) for a dense occupancy grid `occupancy` [res, res, res] (bool / uint8, indexed [x, y, z]) over
    [-1, 1]^3. Nuggets are packed ray after ray and sorted by depth."""
    lib = _lib.load()
    origins, dirs = _lib._f32c(origins, "origins"), _lib._f32c(dirs, "dirs")
    if not occupancy.is_cuda:
        raise _lib.ShaciraError(_lib.ERR_INVALID_ARGUMENT, "occupancy must be a CUDA tensor (no CPU fallback)")
    occ = occupancy.to(torch.uint8).contiguous()
    res, R, dev = occ.shape[0], origins.shape[0], origins.device
    if occ.dim() != 3 or occ.shape[1] != res or occ.shape[2] != res:
        raise _lib.ShaciraError(_lib.ERR_INVALID_ARGUMENT, "occupancy must be [res, res, res]")
    count = torch.empty((R,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib._check(lib.shacira_raytrace_dense_count(_lib._ptr(occ), res, _lib._ptr(origins), _lib._ptr(dirs), R,
                                                     _lib._ptr(count), _lib._stream()))
        incl = torch.cumsum(count.to(torch.int64), 0)
        M = int(incl[-1].item()) if R else 0          # the one host sync of the path: output sizes
        offset = (incl - count).contiguous()
        ridx = torch.empty((M,), dtype=torch.int32, device=dev)
        pidx = torch.empty((M,), dtype=torch.int32, device=dev)
        depth = torch.empty((M, 2), dtype=torch.float32, device=dev)
        if M:
            _lib._check(lib.shacira_raytrace_dense_fill(_lib._ptr(occ), res, _lib._ptr(origins), _lib._ptr(dirs), R,
                                                        _lib._ptr(offset), _lib._ptr(ridx), _lib._ptr(pidx),
                                                        _lib._ptr(depth), _lib._stream()))
    return ridx, pidx, depth


def raymarch_voxel(occupancy, origins, dirs, num_samples, jitter=None):
    """`OctreeAS._raymarch_voxel` (octree_as.py:172-228) on a dense occupancy grid: intersect, then `num_samples`
    stratified samples per intersected cell. Returns (ridx, samples, depth_samples, deltas, boundary)."""
    ridx, _, depth = raytrace_dense(occupancy, origins, dirs)
    return voxel_samples(origins, dirs, ridx, depth, num_samples, jitter)


def prune_dense(occupancy, density_fn, density_decay, min_density, jitter=None):
    """`NeuralRadianceField.prune` (nerf.py:150-185) on a dense grid: `occupancy` float [res, res, res] is the running
    estimate (updated in place: max(density, occupancy * decay)); `density_fn(samples [res^3, 3]) -> [res^3]` is the
    caller's network evaluated at one jittered sample per cell. Returns the uint8 mask grid [res, res, res] that
    `raytrace_dense` / `raymarch_voxel` take (unchanged occupancy when nothing would survive, as the reference)."""
    lib = _lib.load()
    occ = _lib._f32c(occupancy, "occupancy")
    if occ.data_ptr() != occupancy.data_ptr():
        raise _lib.ShaciraError(_lib.ERR_INVALID_ARGUMENT, "occupancy must be contiguous float32 (it is updated in place)")
    res, dev = occ.shape[0], occ.device
    cells = res ** 3
    if jitter is None:
        jitter = torch.rand((cells, 3), dtype=torch.float32, device=dev)
    jitter = _lib._f32c(jitter, "jitter")
    samples = torch.empty((cells, 3), dtype=torch.float32, device=dev)
    mask = torch.empty((res, res, res), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib._check(lib.shacira_prune_samples(res, _lib._ptr(jitter), _lib._ptr(samples), _lib._stream()))
        with torch.no_grad():
            density = _lib._f32c(density_fn(samples).reshape(-1).float(), "density")
        _lib._check(lib.shacira_prune_update(cells, _lib._ptr(density), float(density_decay), float(min_density),
                                             _lib._ptr(occ), _lib._ptr(mask), _lib._stream()))
    return mask
