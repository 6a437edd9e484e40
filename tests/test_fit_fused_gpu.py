"""The tile-resident fused kernel of the image-fit step (csrc/fit_kernels.cuh, shacira_fit_tile_step: grid forward +
decoder MLP + MSE + grid backward in one launch) against
  * the three-kernel path it replaces (tiled forward -> tensor-core MLP -> tiled backward), and
  * plain PyTorch autograd in float64 over the package's own interpolation (grid_ops.latent_hashgrid is pinned to the
    reference by tests/test_latent_gpu.py),
on the Kodak shape (BASELINE cfg2) with STE rounding and with an SGA-style unrounded table. Tolerances: loss 1e-5
relative, every gradient 1e-4 relative per level (north_star: gradients within 1e-4)."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import level_rel_err, level_rms_err, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

BWD_TOL = 1e-4


def _setup(dev, seed, h=512, w=768, max_res=512, random_coords=False):
    import fit_image
    from shacira_b200.grids import LatentGrid
    torch.manual_seed(seed)
    grid = LatentGrid.from_geometric(feature_dim=1, num_lods=16, latent_dim=1, multiscale_type="cat", resolution_dim=2,
                                     feature_std=0.1, codebook_bitwidth=16, min_grid_res=16, max_grid_res=max_res,
                                     init_grid="uniform", conf_latent_decoder=dict(fit_image.DEC),
                                     conf_entropy_reg=dict(fit_image.ENT))
    mlp = torch.nn.Sequential(torch.nn.Linear(16, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                              torch.nn.Linear(16, 3))
    with torch.no_grad():
        grid.codebook.mul_(fit_image.LATENT_SCALE)
        grid.latent_dec.layers[0].shift.normal_(0, 0.05)
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    coords = torch.stack([(ys.reshape(-1) + 0.5) / h * 2 - 1, (xs.reshape(-1) + 0.5) / w * 2 - 1], 1).float()
    if random_coords:   # uneven tiles: their last pass is partial
        coords = torch.rand(h * w, 2) * 2 - 1
    target = torch.rand(h * w, 3)
    return grid.to(dev), mlp.to(dev), coords.to(dev), target.to(dev)


def _run(fs, fused, lat, rflag):
    """One grid + MLP pass of an ImageFitStep's buffers through the fused kernel or the three-kernel path; returns
    (sse, packed MLP gradients, table gradient, sum of dA rows, sum of dshift rows)."""
    from shacira_b200 import _lib
    lib, P, chk = _lib.load(), _lib._ptr, _lib._check
    lin, L = fs.lin, fs.L
    shift = fs.layer.shift.data
    fs.g_grid.zero_()
    fs.g_dec.zero_()
    st = _lib._stream()
    if fused:
        chk(lib.shacira_fit_tile_step(fs.plan.handle, P(lat), fs.fi, fs.rs, L, fs.bw, rflag, P(fs.A), P(shift),
                                      P(fs.target), P(lin[0].weight.data), P(lin[0].bias.data), P(lin[1].weight.data),
                                      P(lin[1].bias.data), P(lin[2].weight.data), P(lin[2].bias.data), fs.T,
                                      P(fs.g_grid), P(fs.g_dec), P(fs.g_dec[L:]), P(fs.mlp_out), st))
    else:
        chk(lib.shacira_latent_forward_planned(fs.plan.handle, P(lat), fs.fi, fs.rs, L, fs.bw, 1, 1, rflag, P(fs.A),
                                               P(shift), 0, P(fs.feats), st))
        chk(lib.shacira_mlp_mse_step_bounded(P(fs.feats), P(fs.target), fs.n, 16, 16, 3, P(lin[0].weight.data),
                                             P(lin[0].bias.data), P(lin[1].weight.data), P(lin[1].bias.data),
                                             P(lin[2].weight.data), P(lin[2].bias.data), P(fs.gfeat), None,
                                             P(fs.mlp_out), P(fs.gfeat_max), st))
        chk(lib.shacira_latent_backward_planned_bounded(fs.plan.handle, P(fs.gfeat), P(lat), fs.fi, fs.rs, L, fs.bw, 1,
                                                        1, rflag, P(fs.A), 0, fs.T, 0, P(fs.g_grid), P(fs.g_dec),
                                                        P(fs.g_dec[L:]), None, st))
    torch.cuda.synchronize()
    n_par = 16 * 16 + 16 + 16 * 16 + 16 + 48 + 3
    sse = float(fs.mlp_out[:2].view(torch.float64)[0])
    return (sse, fs.mlp_out[2:2 + n_par].cpu().numpy().copy(), fs.g_grid.cpu().numpy().copy().reshape(-1),
            float(fs.g_dec[:L].double().sum()), float(fs.g_dec[L:].double().sum()))


@pytest.mark.parametrize("mode", ["ste", "unrounded", "ste-random-coords"])
def test_fused_tile_step_matches_three_kernels_and_fp64(lib, mode):
    from shacira_b200.image_fit import ImageFitStep
    dev = torch.device("cuda", 0)
    grid, mlp, coords, target = _setup(dev, 11, random_coords=mode.endswith("random-coords"))
    mode = mode.split("-")[0]
    fs = ImageFitStep(grid, mlp, coords, target)
    assert fs.fused
    lat = grid.codebook.data if mode == "ste" else (grid.codebook.data + 0.37).contiguous()
    rflag = 1 if mode == "ste" else 0
    _run(fs, True, lat, rflag)                          # (the first call also builds the plan's node table)
    launches0 = lib.launch_count()
    got = _run(fs, True, lat, rflag)
    assert lib.launch_count() - launches0 == 1          # ONE kernel (plus the memset of the 2.4 KB result block)
    ref = _run(fs, False, lat, rflag)
    first = [int(v) for v in grid._first_idx()]
    sizes = [b - a for a, b in zip(first, first[1:] + [fs.T])]
    # against the three-kernel path: the same feature rows (bit-identical arithmetic), the same MLP math
    assert abs(got[0] - ref[0]) <= 1e-6 * abs(ref[0])
    assert rel_err(got[1], ref[1]) <= 2e-5
    assert level_rel_err(got[2], ref[2], first, sizes) <= BWD_TOL
    assert level_rms_err(got[2], ref[2], first, sizes) <= BWD_TOL
    assert abs(got[3] - ref[3]) <= BWD_TOL * max(abs(ref[3]), 1e-6) and abs(got[4] - ref[4]) <= BWD_TOL * max(abs(ref[4]), 1e-6)

    # against float64 autograd: feats = interp(q) * A + shift, MLP, mean squared error
    q = (torch.round(lat) if rflag else lat).double().requires_grad_(True)
    A64 = fs.A.double().clone().requires_grad_(True)
    S64 = fs.layer.shift.data.double().clone().requires_grad_(True)
    csort = coords[fs.plan.perm_tensor().long()]
    # plain torch bilinear gather in float64 (small and explicit: this is the checker, not the product)
    feats = []
    for l, res in enumerate(grid.resolutions):
        t = csort.double() * 0.5 + 0.5
        x = (t * res).float().clamp(0.0, float(np.float32(res - 1 - 1e-5)))
        cell = torch.floor(x)
        fr = (x - cell).double()
        cx, cy = cell[:, 0].long(), cell[:, 1].long()
        dense = res * res < 2 ** 16

        def row(ix, iy):
            if dense:
                r = torch.clamp(ix + iy * res, max=min(2 ** 16, res * res) - 1)
            else:
                r = (ix ^ (iy * 2654435761)) & 0xFFFF
            return q[first[l] + r, 0]
        g0, g1, f0, f1 = 1 - fr[:, 0], 1 - fr[:, 1], fr[:, 0], fr[:, 1]
        feats.append(row(cx, cy) * g0 * g1 + row(cx, cy + 1) * g0 * f1 + row(cx + 1, cy) * f0 * g1 +
                     row(cx + 1, cy + 1) * f0 * f1)
    x64 = torch.stack(feats, 1) * A64.reshape(()) + S64.reshape(())
    lin64 = [(m.weight.data.double().clone().requires_grad_(True), m.bias.data.double().clone().requires_grad_(True))
             for m in fs.lin]
    h = torch.relu(x64 @ lin64[0][0].T + lin64[0][1])
    h = torch.relu(h @ lin64[1][0].T + lin64[1][1])
    y = h @ lin64[2][0].T + lin64[2][1]
    sse = ((y - fs.target.double()) ** 2).sum()
    (sse / (fs.n * 3)).backward()
    print("fused vs three kernels: level_rel %.2e rms %.2e | vs fp64: level_rel %.2e rms %.2e, mlp %.2e" % (
        level_rel_err(got[2], ref[2], first, sizes), level_rms_err(got[2], ref[2], first, sizes),
        level_rel_err(got[2], q.grad[:, 0].cpu().numpy(), first, sizes),
        level_rms_err(got[2], q.grad[:, 0].cpu().numpy(), first, sizes),
        rel_err(got[1], torch.cat([t.grad.reshape(-1) for pair in lin64 for t in pair]).cpu().numpy())))
    print("three kernels vs fp64: level_rel %.2e rms %.2e" % (
        level_rel_err(ref[2], q.grad[:, 0].cpu().numpy(), first, sizes),
        level_rms_err(ref[2], q.grad[:, 0].cpu().numpy(), first, sizes)))
    assert abs(got[0] - float(sse.detach())) <= 1e-5 * float(sse.detach())
    packed64 = torch.cat([t.grad.reshape(-1) for pair in lin64 for t in pair]).cpu().numpy()
    assert rel_err(got[1], packed64) <= BWD_TOL
    gq = q.grad[:, 0].cpu().numpy()
    assert level_rel_err(got[2], gq, first, sizes) <= BWD_TOL
    assert level_rms_err(got[2], gq, first, sizes) <= BWD_TOL
    assert abs(got[3] - float(A64.grad.sum())) <= BWD_TOL * abs(float(A64.grad.sum()))
    assert abs(got[4] - float(S64.grad.sum())) <= BWD_TOL * abs(float(S64.grad.sum()))
    fs.close()


def test_fused_tile_step_nonfinite_and_outliers(lib):
    """A target outlier makes one tile's gradient 2^20 times larger than its neighbours' halfway through the tile
    (exercises the running-maximum rescale of the fixed-point accumulators); an Inf target poisons the touched nodes
    with NaN like float atomics would and leaves the rest of the table finite."""
    from shacira_b200.image_fit import ImageFitStep
    dev = torch.device("cuda", 0)
    grid, mlp, coords, target = _setup(dev, 12)
    fs = ImageFitStep(grid, mlp, coords, target)
    lat = grid.codebook.data
    n = fs.n
    fs.target[n // 2 + 300, 1] = 3.0e6          # sorted order: somewhere inside a tile, not in its first pass
    got = _run(fs, True, lat, 1)
    ref = _run(fs, False, lat, 1)
    f = list(grid._first_idx()) + [fs.T]
    sizes = [f[i + 1] - f[i] for i in range(16)]
    assert np.isfinite(got[2]).all()
    assert level_rel_err(got[2], ref[2], grid._first_idx(), sizes) <= BWD_TOL
    fs.target[n // 2 + 300, 1] = float("inf")
    got = _run(fs, True, lat, 1)
    bad = ~np.isfinite(got[2])
    assert bad.any() and bad.sum() <= 4 * 16 * 1300   # only nodes of the outlier's tile
    fs.close()


def test_fit_step_fused_matches_unfused_after_steps(lib, monkeypatch):
    """Whole ImageFitStep (SGA phase and STE phase, injected noise): parameters after 5 steps with the fused kernel
    against the three-kernel path."""
    import copy
    from shacira_b200.image_fit import ImageFitStep
    dev = torch.device("cuda", 0)
    grid, mlp, coords, target = _setup(dev, 13)
    grid2, mlp2 = copy.deepcopy(grid), copy.deepcopy(mlp)
    fa = ImageFitStep(grid, mlp, coords, target)
    monkeypatch.setenv("SHACIRA_FIT_FUSED", "0")
    fb = ImageFitStep(grid2, mlp2, coords, target)
    monkeypatch.delenv("SHACIRA_FIT_FUSED")
    assert fa.fused and not fb.fused
    gen = torch.Generator().manual_seed(7)
    T = fa.T
    for it in range(5):
        noise = (torch.rand((T, 1), generator=gen) - 0.5).to(dev)
        u = torch.rand((T, 1, 2), generator=gen).to(dev)
        for f in (fa, fb):
            f.set_sga(it < 3)
            f.set_temperature(0.5)
            f.sga_uniforms = u
            f.set_lambda(1e-3)
            f.noise.copy_(noise)
            f.step()
        assert abs(float(fa.rgb_loss()) - float(fb.rgb_loss())) <= 1e-4 * float(fb.rgb_loss())
    d = (grid.codebook.data - grid2.codebook.data).abs()
    assert float((d <= 1e-3 * float(grid2.codebook.data.abs().max())).float().mean()) >= 0.999
    for p, q in zip(mlp.parameters(), mlp2.parameters()):
        assert float((p.data - q.data).abs().max()) <= 2e-3 * float(q.data.abs().max())
    fa.close()
    fb.close()


def test_one_launch_optimizer_matches_separate_kernels(lib, monkeypatch):
    """shacira_fit_optimizer_step (bit-rate loss + small tensors + table Adam + the NEXT step's SGA sample in one launch)
    against the separate bit-rate / SGA / table-Adam / small-tensor-Adam launches: the draws are counter based (same index -> same sample), so
    six steps (four with SGA, two with STE rounding) end in the same parameters."""
    import copy
    from shacira_b200.image_fit import ImageFitStep
    dev = torch.device("cuda", 0)
    grid, mlp, coords, target = _setup(dev, 14)
    grid2, mlp2 = copy.deepcopy(grid), copy.deepcopy(mlp)
    fa = ImageFitStep(grid, mlp, coords, target, device_noise=True, noise_seed=3)
    monkeypatch.setenv("SHACIRA_FIT_OPT_FUSED", "0")
    fb = ImageFitStep(grid2, mlp2, coords, target, device_noise=True, noise_seed=3)
    monkeypatch.delenv("SHACIRA_FIT_OPT_FUSED")
    assert fa.opt_fused and not fb.opt_fused
    for it in range(6):
        for f in (fa, fb):
            if it == 0:
                f.set_sga(True)
                f.set_temperature(0.7)
            if it == 4:
                f.set_sga(False)
            f.set_lambda(1e-3)
        n0 = lib.launch_count()
        fa.step()
        na = lib.launch_count() - n0
        fb.step()
        nb = lib.launch_count() - n0 - na
        if 0 < it < 4:
            assert na == 2 and nb == 6, (na, nb)      # fused tile kernel + optimizer (bit-rate loss, all Adams, next SGA sample inside)
                                                     # against bit-rate, SGA, its counter, tile kernel, two Adam launches
        assert abs(float(fa.rgb_loss()) - float(fb.rgb_loss())) <= 1e-5 * float(fb.rgb_loss())
        assert abs(float(fa.total_bits()) - float(fb.total_bits())) <= 1e-5 * float(fb.total_bits())
    assert float((grid.codebook.data - grid2.codebook.data).abs().max()) <= 1e-5 * float(grid2.codebook.data.abs().max())
    for p, q in zip(mlp.parameters(), mlp2.parameters()):
        assert float((p.data - q.data).abs().max()) <= 1e-5 * float(q.data.abs().max())
    # (the one-launch form has drawn one sample ahead -- discarded when SGA was switched off)
    assert int(fa.sga_rng_step) == int(fb.sga_rng_step) + 1
    fa.close()
    fb.close()
