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
