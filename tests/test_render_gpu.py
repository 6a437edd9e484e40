"""Packed exponential integration against the float64 restatement of the published kaolin algorithm
(oracle/render_oracle.py; kaolin itself is absent: parity with the package is unpinned)."""
import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _packed_case(num_rays, max_len, nf, seed, min_len=1):
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len + 1, size=num_rays)
    S = int(lens.sum())
    boundary = np.zeros(S, dtype=bool)
    boundary[np.concatenate(([0], np.cumsum(lens)[:-1]))] = True
    feats = rng.random((S, nf)).astype(np.float32)
    tau = (rng.random(S) ** 3 * 4.0).astype(np.float32)     # mostly thin samples, a few opaque ones
    return feats, tau, boundary


@pytest.mark.parametrize("num_rays,max_len,nf", [(1, 1, 3), (7, 31, 3), (64, 128, 3), (300, 200, 4), (50, 33, 1),
                                                 (4096, 128, 3), (20, 1000, 8)])
def test_integration_forward_and_backward(lib, num_rays, max_len, nf):
    from oracle import render_oracle as ro
    from shacira_b200 import render
    feats, tau, boundary = _packed_case(num_rays, max_len, nf, seed=num_rays + nf)
    f = torch.from_numpy(feats).cuda().requires_grad_(True)
    t = torch.from_numpy(tau).cuda().unsqueeze(1).requires_grad_(True)
    b = torch.from_numpy(boundary).cuda()
    ray, w = render.exponential_integration(f, t, b, exclusive=True)
    alpha = render.sum_reduce(w, b)
    want_ray, want_w = ro.exponential_integration(feats, tau, boundary)
    assert ray.shape == (num_rays, nf) and w.shape == (tau.shape[0], 1)
    assert rel_err(ray.detach().cpu().numpy(), want_ray) <= 1e-5
    assert rel_err(w.detach().cpu().numpy().reshape(-1), want_w) <= 1e-5
    starts = list(np.nonzero(boundary)[0]) + [tau.shape[0]]
    want_alpha = np.array([want_w[starts[r]:starts[r + 1]].sum() for r in range(num_rays)])
    assert rel_err(alpha.detach().cpu().numpy().reshape(-1), want_alpha) <= 1e-5
    # gradients: a loss that uses the ray colours, alpha and a depth-like weighted sum, as the tracer does
    torch.manual_seed(0)
    g_ray = torch.randn(num_rays, nf, device="cuda")
    depth = torch.rand(tau.shape[0], 1, device="cuda")
    loss = (ray * g_ray).sum() + (alpha ** 2).sum() + render.sum_reduce(depth * w, b).sum()
    loss.backward()
    f64 = torch.from_numpy(feats).double().requires_grad_(True)
    t64 = torch.from_numpy(tau).double().requires_grad_(True)
    ray64, w64 = ro.exponential_integration_torch(f64, t64, torch.from_numpy(boundary))
    seg = torch.from_numpy(np.repeat(np.arange(num_rays), np.diff(starts)))
    alpha64 = torch.zeros(num_rays, dtype=torch.float64).index_add_(0, seg, w64)
    loss64 = (ray64 * g_ray.cpu().double()).sum() + (alpha64 ** 2).sum() + (depth.cpu().double().reshape(-1) * w64).sum()
    loss64.backward()
    assert rel_err(f.grad.cpu().numpy(), f64.grad.numpy()) <= 1e-4
    assert rel_err(t.grad.cpu().numpy().reshape(-1), t64.grad.numpy()) <= 1e-4


def test_integration_rejects_cpu_tensors_and_odd_widths(lib):
    from shacira_b200 import render
    b = torch.tensor([True, False, False])
    with pytest.raises(lib.ShaciraError):
        render.exponential_integration(torch.zeros(3, 3), torch.zeros(3, 1), b)
    with pytest.raises(lib.ShaciraError):
        render.exponential_integration(torch.zeros(3, 5).cuda(), torch.zeros(3, 1).cuda(), b.cuda())


def test_voxel_samples_match_reference_helpers_bit_exactly(lib):
    """The fused sample-generation kernel against the golden outputs of the reference's own sampling.py
    (depth samples, boundary: bit-exact) and the oracle restatement (deltas bit-exact, samples to 1 ulp-ish: the
    reference's addcmul may or may not contract to an FMA)."""
    import os
    from oracle import render_oracle as ro
    from shacira_b200 import render
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampling_ref.npz"))
    for name in ("a", "b", "c", "d"):
        depth, jitter, ridx = g[name + "/depth"], g[name + "/jitter"], g[name + "/ridx"]
        K = int(g[name + "/K"])
        R = int(ridx.max()) + 1
        rng = np.random.default_rng(1)
        o, d = rng.standard_normal((R, 3)).astype(np.float32), rng.standard_normal((R, 3)).astype(np.float32)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        ridx_out, samples, ds, deltas, boundary = render.voxel_samples(dev(o), dev(d), dev(ridx), dev(depth), K, dev(jitter))
        want = ro.voxel_samples(o, d, ridx, depth, jitter)
        assert np.array_equal(ds.cpu().numpy().reshape(-1, K).view(np.uint32), g[name + "/depth_samples"].view(np.uint32))
        assert np.array_equal(boundary.cpu().numpy().astype(np.uint8), g[name + "/boundary"])
        assert np.array_equal(ridx_out.cpu().numpy(), want[0])
        assert np.array_equal(deltas.cpu().numpy().reshape(-1).view(np.uint32), want[3].view(np.uint32))
        assert np.allclose(samples.cpu().numpy(), want[1], rtol=1e-6, atol=1e-6)
    # device-drawn jitter: stratified, inside the intervals, boundaries usable by the integration
    M, K = 1000, 16
    depth = torch.rand(M, 1, device="cuda") * 2
    depth = torch.cat((depth, depth + 0.1), 1)
    ridx = torch.sort(torch.randint(0, 100, (M,), device="cuda"))[0]
    o, d = torch.randn(100, 3, device="cuda"), torch.randn(100, 3, device="cuda")
    r2, s2, ds2, dl2, b2 = render.voxel_samples(o, d, ridx, depth, K)
    ds2 = ds2.reshape(M, K)
    assert bool((ds2 >= depth[:, :1]).all()) and bool((ds2 <= depth[:, 1:]).all()) and bool((ds2.diff(dim=1) > 0).all())
    assert int(b2.sum()) == int(torch.unique(ridx).numel()) and bool((dl2 >= 0).all())


@pytest.mark.parametrize("res,fill", [(8, 0.3), (16, 0.1), (4, 1.0), (32, 0.02)])
def test_dense_raytrace_matches_bruteforce_slab_tests(lib, res, fill):
    """3D-DDA kernel (stand-in for kaolin's unbatched_raytrace on the dense grid) against slab tests of every occupied
    cell. Grazing hits (interval shorter than 1e-5) may legitimately differ and are ignored on both sides."""
    from oracle import render_oracle as ro
    from shacira_b200 import render
    rng = np.random.default_rng(res)
    occ = rng.random((res, res, res)) < fill
    R = 200
    o = (rng.standard_normal((R, 3)) * 1.5).astype(np.float32)
    o[: R // 4] = (rng.random((R // 4, 3)) * 1.6 - 0.8).astype(np.float32)       # some origins inside the box
    tgt = (rng.random((R, 3)) * 2 - 1).astype(np.float32)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[5] = [1.0, 0.0, 0.0]                                                        # axis-parallel rays
    d[6] = [0.0, -1.0, 0.0]
    d = d.astype(np.float32)
    ridx, pidx, depth = render.raytrace_dense(torch.from_numpy(occ).cuda(), torch.from_numpy(o).cuda(),
                                              torch.from_numpy(d).cuda())
    ridx, pidx, depth = ridx.cpu().numpy(), pidx.cpu().numpy(), depth.cpu().numpy()
    assert np.all(np.diff(ridx) >= 0)                                              # packed ray after ray
    want = ro.raytrace_dense_bruteforce(occ, o, d)
    eps = 1e-5
    total = 0
    for r in range(R):
        got = [(int(c), float(a), float(b)) for c, (a, b) in zip(pidx[ridx == r], depth[ridx == r]) if b - a > eps]
        ref = [h for h in want[r] if h[2] - h[1] > eps]
        assert [g[0] for g in got] == [h[0] for h in ref], r
        for g, h in zip(got, ref):
            assert abs(g[1] - h[1]) <= 1e-4 * max(1.0, h[1]) and abs(g[2] - h[2]) <= 1e-4 * max(1.0, h[2])
        ent = [g[1] for g in got]
        assert ent == sorted(ent)                                                  # sorted by depth along the ray
        total += len(got)
    assert total > 0
    # end to end: the reference's voxel raymarch on the dense grid, then integrate
    r2, samples, ds, deltas, boundary = render.raymarch_voxel(torch.from_numpy(occ).cuda(), torch.from_numpy(o).cuda(),
                                                              torch.from_numpy(d).cuda(), 4)
    assert samples.shape[0] == 4 * pidx.shape[0] and bool((samples.abs() <= 1.0 + 1e-4).all())
    if samples.shape[0]:
        ray, w = render.exponential_integration(torch.rand(samples.shape[0], 3, device="cuda"), deltas * 5.0, boundary)
        assert ray.shape[0] == int(boundary.sum()) and bool(torch.isfinite(ray).all())


def test_prune_dense_matches_reference_restatement(lib):
    """NeuralRadianceField.prune on the dense grid (nerf.py:158-171): samples, running occupancy and mask bit-exact
    against the numpy restatement with the same jitter and densities; the mask drives the ray tracer."""
    from oracle import render_oracle as ro
    from shacira_b200 import render
    res = 16
    rng = np.random.default_rng(3)
    occ0 = rng.random(res ** 3).astype(np.float32) * 0.05
    jitter = rng.random((res ** 3, 3)).astype(np.float32)
    dens = (rng.random(res ** 3).astype(np.float32) ** 4 * 0.2)
    occ = torch.from_numpy(occ0.copy()).cuda().reshape(res, res, res)
    seen = {}

    def density_fn(samples):
        seen["samples"] = samples.clone()
        return torch.from_numpy(dens).cuda()

    mask = render.prune_dense(occ, density_fn, 0.6, 0.01, jitter=torch.from_numpy(jitter).cuda())
    want_s, want_occ, want_mask = ro.prune_dense(occ0, dens, jitter, 0.6, 0.01)
    assert np.array_equal(seen["samples"].cpu().numpy().view(np.uint32), want_s.view(np.uint32))
    assert np.array_equal(occ.cpu().numpy().reshape(-1).view(np.uint32), want_occ.view(np.uint32))
    assert np.array_equal(mask.cpu().numpy().reshape(-1).astype(bool), want_mask) and 0 < int(mask.sum()) < res ** 3
    o = torch.tensor([[-3.0, 0.1, 0.2]], device="cuda")
    d = torch.tensor([[1.0, 0.0, 0.0]], device="cuda")
    ridx, pidx, depth = render.raytrace_dense(mask, o, d)
    assert bool(mask.reshape(-1)[pidx.long()].all())             # only surviving cells are hit


def test_render_helpers_on_empty_and_missing_inputs(lib):
    """Rays that miss the grid, zero rays, zero nuggets: empty outputs, no launches on NULL pointers."""
    from shacira_b200 import render
    occ = torch.ones((4, 4, 4), dtype=torch.uint8, device="cuda")
    o = torch.tensor([[5.0, 5.0, 5.0], [0.0, 3.0, 0.0]], device="cuda")
    d = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], device="cuda")       # both point away from the box
    ridx, pidx, depth = render.raytrace_dense(occ, o, d)
    assert ridx.numel() == 0 and pidx.numel() == 0 and depth.shape == (0, 2)
    r2, samples, ds, deltas, boundary = render.raymarch_voxel(occ, o, d, 8)
    assert samples.shape == (0, 3) and ds.shape == (0, 1) and boundary.numel() == 0
    ray, w = render.exponential_integration(torch.zeros((0, 3), device="cuda"), torch.zeros((0, 1), device="cuda"),
                                            torch.zeros((0,), dtype=torch.bool, device="cuda"))
    assert ray.shape == (0, 3) and w.shape == (0, 1)
    ridx0, _, _ = render.raytrace_dense(occ, torch.zeros((0, 3), device="cuda"), torch.zeros((0, 3), device="cuda"))
    assert ridx0.numel() == 0
    # an empty grid: nothing is hit even by rays through the box
    none = torch.zeros((4, 4, 4), dtype=torch.uint8, device="cuda")
    ridx1, _, _ = render.raytrace_dense(none, torch.tensor([[-3.0, 0.1, 0.1]], device="cuda"),
                                        torch.tensor([[1.0, 0.0, 0.0]], device="cuda"))
    assert ridx1.numel() == 0
