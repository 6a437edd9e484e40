"""The tiled fast path (spatial plan + shared-memory node staging + fixed-point accumulation)
against the point-parallel kernels, the CPU oracle and the golden reference vectors."""
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import rel_err

pytestmark = pytest.mark.gpu
FWD_TOL, BWD_TOL = 1e-5, 1e-4


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _case(dim, L, bw, rmin, rmax, n, C, F, seed, per_level=False, kind="uniform"):
    rng = np.random.default_rng(seed)
    res = oracle.geometric_resolutions(rmin, rmax, L)
    sizes, first, T = oracle.level_layout(res, bw, dim)
    if kind == "uniform":
        coords = (rng.random((n, dim), dtype=np.float32) * 2 - 1)
    elif kind == "arbitrary":
        coords = np.clip(rng.standard_normal((n, dim)) * 0.5, -1.2, 1.2).astype(np.float32)
    elif kind == "clustered":  # every point inside one tile: exercises the batch loop of the backward
        coords = (rng.random((n, dim), dtype=np.float32) * 0.01 + 0.3).astype(np.float32)
    elif kind == "pixels":
        H, W = 512, 768
        ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        coords = np.stack([(ys.reshape(-1) / H - 0.5) * 2, (xs.reshape(-1) / W - 0.5) * 2], 1).astype(np.float32)
        coords = coords[rng.permutation(coords.shape[0])][:n]
    lat = ((rng.random((T, C), dtype=np.float32) - 0.5) * 16).astype(np.float32)
    nA = L if per_level else 1
    A = (rng.standard_normal((nA, C, F)) * 0.3).astype(np.float32)
    S = (rng.standard_normal((nA, F)) * 0.1).astype(np.float32)
    g = rng.standard_normal((coords.shape[0], L * F)).astype(np.float32)
    return dict(dim=dim, L=L, bw=bw, res=res, first=first, T=T, coords=np.ascontiguousarray(coords), lat=lat, A=A, S=S,
                g=g, C=C, F=F, nA=nA)


def _oracle_fwd_bwd(c, round_flag=True):
    """Reference order on the CPU: decode table (float64 affine on rounded latents is exact enough as a
    yardstick), interpolate with the C oracle; gradients by the oracle's backward + chain rule."""
    q = np.rint(c["lat"]) if round_flag else c["lat"]
    L, C, F = c["L"], c["C"], c["F"]
    z = oracle.forward(c["coords"], q.astype(np.float32), c["first"], c["res"], c["bw"]).reshape(-1, L, C)
    feats = np.stack([z[:, l].astype(np.float64) @ c["A"][l if c["nA"] > 1 else 0] + c["S"][l if c["nA"] > 1 else 0]
                      for l in range(L)], 1).reshape(-1, L * F)
    g = c["g"].reshape(-1, L, F).astype(np.float64)
    gz = np.stack([g[:, l] @ c["A"][l if c["nA"] > 1 else 0].T for l in range(L)], 1)  # [n, L, C]
    gl = oracle.backward(c["coords"], gz.reshape(-1, L * C).astype(np.float32), c["T"], c["first"], c["res"], c["bw"], C)
    gA = np.einsum("nlc,nlf->lcf", z.astype(np.float64), g)
    gS = g.sum(0)
    return feats, gl, gA, gS


def _check_decoder_grads(c, gA, gS, want_gA, want_gS):
    """Per-level decoders: every level's gradient. One shared decoder: only the sum over levels is defined
    (include/shacira_b200.h) -- the tiled kernel reports it in level slot 0."""
    gA, gS = gA.cpu().numpy().astype(np.float64), gS.cpu().numpy().astype(np.float64)
    if c["nA"] == 1:
        gA, gS, want_gA, want_gS = gA.sum(0), gS.sum(0), want_gA.sum(0), want_gS.sum(0)
    assert rel_err(gA, want_gA) <= BWD_TOL
    assert rel_err(gS, want_gS) <= BWD_TOL


TILED_CASES = [
    # dim, L, bw, rmin, rmax, n, C, F, per_level, kind
    (2, 16, 16, 16, 512, 768 * 512, 1, 1, False, "pixels"),   # BASELINE cfg2 (full size)
    (2, 16, 14, 16, 512, 1 << 18, 1, 1, False, "uniform"),    # BASELINE cfg1 grid
    (2, 16, 16, 16, 512, 40000, 2, 4, True, "arbitrary"),
    (2, 24, 11, 16, 512, 50000, 1, 1, False, "uniform"),      # the reference's kodak.yaml grid: 24 levels, 2^11
    (2, 8, 19, 16, 700, 30000, 4, 2, False, "uniform"),       # dense levels with res >= 257 (Q4)
    (2, 16, 16, 16, 512, 20000, 1, 1, False, "clustered"),
    (3, 16, 19, 16, 2048, 1 << 17, 1, 4, False, "uniform"),   # BASELINE cfg4 grid: fine levels fall back to direct
    (3, 8, 14, 8, 128, 30000, 2, 2, True, "arbitrary"),
]


@pytest.mark.parametrize("dim,L,bw,rmin,rmax,n,C,F,per_level,kind", TILED_CASES)
def test_tiled_forward_backward(lib, dim, L, bw, rmin, rmax, n, C, F, per_level, kind):
    c = _case(dim, L, bw, rmin, rmax, n, C, F, seed=dim * 7 + C + F, per_level=per_level, kind=kind)
    coords, lat, A, S, g = _dev(c["coords"]), _dev(c["lat"]), _dev(c["A"]), _dev(c["S"]), _dev(c["g"])
    plan = lib.Plan(coords)
    info = plan.info()
    assert info["n"] == coords.shape[0] and info["ntiles"] == info["tiles_per_axis"] ** dim
    feats = lib.latent_forward_planned(plan, lat, c["first"], c["res"], bw, A, S, F, True)
    # bit-identical to the point-parallel kernel: same indices, weights, rounding and operation order
    feats_pp, z_pp = lib.latent_forward(coords, lat, c["first"], c["res"], bw, A, S, F, True, True)
    assert torch.equal(feats, feats_pp)
    want_f, want_gl, want_gA, want_gS = _oracle_fwd_bwd(c)
    assert rel_err(feats.cpu().numpy(), want_f) <= FWD_TOL
    gl, gA, gS = lib.latent_backward_planned(plan, g, lat, c["first"], c["res"], bw, A, C, F, c["T"], True, True)
    assert rel_err(gl.cpu().numpy(), want_gl) <= BWD_TOL
    _check_decoder_grads(c, gA, gS, want_gA, want_gS)
    # without decoder gradients (frozen decoder): same latent gradient
    gl2, _, _ = lib.latent_backward_planned(plan, g, None, c["first"], c["res"], bw, A, C, F, c["T"], True, False)
    assert rel_err(gl2.cpu().numpy(), want_gl) <= BWD_TOL
    plan.close()


def test_plan_is_a_valid_spatial_binning(lib):
    c = _case(2, 4, 10, 8, 64, 50000, 1, 1, seed=3, kind="arbitrary")
    coords = _dev(c["coords"])
    plan = lib.Plan(coords)
    perm, cs, off = plan.arrays()
    info = plan.info()
    g = info["tiles_per_axis"]
    assert sorted(perm.tolist()) == list(range(50000))               # a permutation
    assert np.array_equal(cs, c["coords"][perm])                      # sorted copy of the coordinates
    assert off[0] == 0 and off[-1] == 50000 and np.all(np.diff(off) >= 0)
    t = np.clip(np.floor((c["coords"].astype(np.float64) * 0.5 + 0.5) * g), 0, g - 1).astype(np.int64)
    tile = t[:, 0] + g * t[:, 1]
    seg = np.repeat(np.arange(info["ntiles"]), np.diff(off))
    assert np.array_equal(tile[perm], seg)                            # every point sits in its tile's segment
    plan.close()


def test_direct_level_fallback_matches(lib, monkeypatch):
    """Shrinking the shared-memory budget pushes fine levels to the in-kernel direct path."""
    c = _case(2, 16, 16, 16, 512, 30000, 1, 1, seed=9, kind="uniform")
    coords, lat, A, S, g = _dev(c["coords"]), _dev(c["lat"]), _dev(c["A"]), _dev(c["S"]), _dev(c["g"])
    plan = lib.Plan(coords)
    want_f, want_gl, want_gA, want_gS = _oracle_fwd_bwd(c)
    full = lib.latent_forward_planned(plan, lat, c["first"], c["res"], 16, A, S, 1, True)
    for budget in ("1024", "4096"):
        monkeypatch.setenv("SHACIRA_TILE_SMEM", budget)
        f = lib.latent_forward_planned(plan, lat, c["first"], c["res"], 16, A, S, 1, True)
        assert torch.equal(f, full)
        gl, gA, gS = lib.latent_backward_planned(plan, g, lat, c["first"], c["res"], 16, A, 1, 1, c["T"], True, True)
        assert rel_err(gl.cpu().numpy(), want_gl) <= BWD_TOL
        _check_decoder_grads(c, gA, gS, want_gA, want_gS)
    plan.close()


def test_tiled_backward_is_deterministic_per_plan(lib):
    """Fixed-point shared-memory sums are order-independent; only the few per-node flush adds are float."""
    c = _case(2, 16, 16, 16, 512, 60000, 1, 1, seed=4, kind="uniform")
    coords, lat, A, g = _dev(c["coords"]), _dev(c["lat"]), _dev(c["A"]), _dev(c["g"])
    plan = lib.Plan(coords)
    a, _, _ = lib.latent_backward_planned(plan, g, None, c["first"], c["res"], 16, A, 1, 1, c["T"], True, False)
    b, _, _ = lib.latent_backward_planned(plan, g, None, c["first"], c["res"], 16, A, 1, 1, c["T"], True, False)
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-6
    plan.close()


def test_gradient_magnitude_extremes(lib):
    """Tiny, huge and zero upstream gradients keep the fixed-point path within tolerance."""
    c = _case(2, 12, 14, 16, 256, 30000, 1, 1, seed=5, kind="uniform")
    coords, lat, A = _dev(c["coords"]), _dev(c["lat"]), _dev(c["A"])
    plan = lib.Plan(coords)
    for scale in (1e-12, 1.0, 1e12):
        c["g"] = (np.random.default_rng(6).standard_normal(c["g"].shape) * scale).astype(np.float32)
        c["g"][:, 3] = 0.0                      # one level with no gradient at all
        c["g"][7, 5] = 50.0 * scale             # one outlier
        _, want_gl, _, _ = _oracle_fwd_bwd(c)
        gl, _, _ = lib.latent_backward_planned(plan, _dev(c["g"]), None, c["first"], c["res"], 14, A, 1, 1, c["T"], True, False)
        assert rel_err(gl.cpu().numpy(), want_gl) <= BWD_TOL
        a, b = c["first"][3], c["first"][4]
        assert not bool(gl[a:b].any())
    plan.close()


@pytest.mark.parametrize("bad", [np.inf, -np.inf, np.nan])
@pytest.mark.parametrize("bounded", [False, True])
def test_non_finite_upstream_gradient_propagates(lib, bad, bounded):
    """Fixed-point accumulation cannot represent Inf/NaN: the affected tile/level is poisoned with NaN, as the
    reference's float atomics would do, instead of silently dropping the gradient. NaN needs care: fmaxf() returns
    the non-NaN operand, so the max pass (and a caller's bound) must carry it explicitly."""
    c = _case(2, 8, 12, 16, 128, 20000, 1, 1, seed=12, kind="uniform")
    coords, lat, A = _dev(c["coords"]), _dev(c["lat"]), _dev(c["A"])
    g = c["g"].copy()
    g[123, 2] = bad
    plan = lib.Plan(coords)
    level_max = _dev(np.abs(g).max(0).astype(np.float32)) if bounded else None   # np max propagates NaN
    gl, _, _ = lib.latent_backward_planned(plan, _dev(g), None, c["first"], c["res"], 12, A, 1, 1, c["T"], True, False,
                                           level_max=level_max)
    a, b = c["first"][2], c["first"][3]
    assert bool(torch.isnan(gl[a:b]).any())                   # level 2 carries the poison
    assert bool(torch.isfinite(gl[:a]).all()) and bool(torch.isfinite(gl[b:]).all())
    plan.close()


def test_latent_grid_api_uses_the_plan_and_matches_unplanned(lib, monkeypatch):
    from shacira_b200 import grid_ops
    from shacira_b200.grids import LatentGrid
    dec = dict(ldecode_enabled=True, ldecode_type="single", use_sga=False, diff_sampling=True, use_shift=True,
               ldecode_matrix="sq", latent_dim=1, norm="max", norm_every=10, ldec_std=0.1, decay_period=0.9,
               temperature=0.1)
    ent = dict(num_prob_layers=2, entropy_reg=1e-3, entropy_reg_end=1e-4, entropy_reg_sched="cosine", noise_freq=1)
    torch.manual_seed(0)
    grid = LatentGrid.from_geometric(feature_dim=1, num_lods=16, latent_dim=1, multiscale_type="cat", resolution_dim=2,
                                     feature_std=0.1, codebook_bitwidth=16, min_grid_res=16, max_grid_res=512,
                                     init_grid="uniform", conf_latent_decoder=dec, conf_entropy_reg=ent)
    with torch.no_grad():
        grid.codebook.mul_(60.0)
    grid = grid.cuda()
    coords = (torch.rand(100000, 2, device="cuda") * 2 - 1)
    gout = torch.randn(100000, 16, device="cuda")
    grid_ops.clear_plans()
    h0, b0 = grid_ops.plan_stats["hits"], grid_ops.plan_stats["builds"]
    outs = []
    for _ in range(3):                                   # static coordinates: one plan, reused
        grid.zero_grad()
        f = grid.interpolate(coords, 0)
        f.backward(gout)
        outs.append((f.detach().clone(), grid.codebook.grad.clone(), grid.latent_dec.layers[0].scale.grad.clone(),
                     grid.latent_dec.layers[0].shift.grad.clone()))
    assert grid_ops.plan_stats["builds"] == b0 + 1 and grid_ops.plan_stats["hits"] >= h0 + 2
    coords.mul_(0.5)                                     # in-place change -> version bump -> new plan
    grid.interpolate(coords, 0)
    assert grid_ops.plan_stats["builds"] == b0 + 2
    coords.mul_(2.0)
    monkeypatch.setenv("SHACIRA_DISABLE_PLAN", "1")
    grid.zero_grad()
    f = grid.interpolate(coords, 0)
    f.backward(gout)
    assert torch.equal(f, outs[0][0])
    assert rel_err(outs[0][1].cpu().numpy(), grid.codebook.grad.cpu().numpy()) <= BWD_TOL
    assert rel_err(outs[0][2].cpu().numpy(), grid.latent_dec.layers[0].scale.grad.cpu().numpy()) <= BWD_TOL
    assert rel_err(outs[0][3].cpu().numpy(), grid.latent_dec.layers[0].shift.grad.cpu().numpy()) <= BWD_TOL


def test_plan_cache_recycles_allocations_for_changing_coordinates(lib, monkeypatch):
    """Fresh coordinates every step (NeRF-like): the cache recycles plan objects instead of allocating."""
    from shacira_b200 import grid_ops
    c = _case(2, 8, 12, 16, 128, 20000, 1, 1, seed=21, kind="uniform")
    lat, A, S = _dev(c["lat"]).requires_grad_(True), _dev(c["A"]), _dev(c["S"])
    grid_ops.clear_plans()
    monkeypatch.setattr(grid_ops, "PLAN_CACHE_SIZE", 2)
    monkeypatch.setattr(grid_ops, "PLAN_MIN_POINTS", 1)     # the mechanism, not the crossover, is under test
    handles = set()
    for step in range(6):
        coords = torch.rand(20000, 2, device="cuda") * 2 - 1
        f = grid_ops.latent_hashgrid(coords, lat, A, S, c["first"], c["res"], 12, True)
        want, _ = lib.latent_forward(coords, lat.detach(), c["first"], c["res"], 12, A, S, 1, True, False)
        assert torch.equal(f.detach(), want)
        f.sum().backward()
        handles.update(p.handle.value for p in grid_ops._plans.values())
    assert len(grid_ops._plans) == 2 and len(handles) == 2   # two plan objects serve all six coordinate sets
    # a plan that a live autograd graph still needs is never recycled under it
    grid_ops.clear_plans()
    pending = []
    for step in range(4):
        coords = torch.rand(20000, 2, device="cuda") * 2 - 1
        pending.append((coords, grid_ops.latent_hashgrid(coords, lat, A, S, c["first"], c["res"], 12, True)))
    lat.grad = None
    pending[0][1].sum().backward()      # its plan left the cache two forwards ago
    g_first = lat.grad.clone()
    grid_ops.clear_plans()
    lat.grad = None
    grid_ops.latent_hashgrid(pending[0][0], lat, A, S, c["first"], c["res"], 12, True).sum().backward()
    assert rel_err(g_first.cpu().numpy(), lat.grad.cpu().numpy()) <= 1e-6
    grid_ops.clear_plans()


@pytest.mark.parametrize("C,F,dec", [(1, 1, True), (1, 1, False), (2, 4, True), (4, 2, False)])
def test_bounded_backward_matches_self_scaled(lib, C, F, dec):
    """shacira_latent_backward_planned_bounded: with the caller's per-column bound of |grad_output| the kernel skips
    its own max pass; the result agrees with the oracle to the same tolerance, also when the bound is loose."""
    c = _case(2, 16, 16, 16, 512, 100000, C, F, seed=91 + C + F, kind="uniform")
    coords, lat, A, g = _dev(c["coords"]), _dev(c["lat"]), _dev(c["A"]), _dev(c["g"])
    # uneven magnitudes across levels and space, like real upstream gradients
    g = g * torch.logspace(-3, 2, g.shape[1], device="cuda") * (0.1 + coords[:, :1].abs())
    c["g"] = g.cpu().numpy()
    plan = lib.Plan(coords)
    _, want_gl, want_gA, want_gS = _oracle_fwd_bwd(c)
    for slack in (1.0, 7.3):
        bound = g.abs().amax(dim=0) * slack
        gl, gA, gS = lib.latent_backward_planned(plan, g, lat if dec else None, c["first"], c["res"], 16, A, C, F, c["T"],
                                                 True, dec, level_max=bound)
        bounds = list(c["first"]) + [c["T"]]
        for l in range(16):   # per level: the coarse levels must not hide behind the largest gradient
            a, b = bounds[l], bounds[l + 1]
            assert rel_err(gl[a:b].cpu().numpy(), want_gl[a:b]) <= BWD_TOL, (l, slack)
        if dec:
            _check_decoder_grads(c, gA, gS, want_gA, want_gS)
    plan.close()


def test_mlp_step_reports_feature_gradient_bound(lib):
    torch.manual_seed(5)
    n = 50000
    mlp = torch.nn.Sequential(torch.nn.Linear(16, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                              torch.nn.Linear(16, 3)).cuda()
    x, gt = torch.randn(n, 16, device="cuda"), torch.rand(n, 3, device="cuda")
    bound = torch.full((16,), -1.0, device="cuda")
    ws = [mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias, mlp[4].weight, mlp[4].bias]
    _, gx, _, _ = lib.mlp_mse_step(x, gt, *[w.detach() for w in ws], absmax_out=bound)
    assert torch.equal(bound, gx.abs().amax(dim=0))


def test_sorted_io_plan_exchanges_rows_in_tile_order(lib):
    """shacira_plan_set_sorted_io: forward rows come out at the sorted position, backward rows are read there;
    values are the same numbers as with the default (original-index) exchange."""
    c = _case(2, 16, 16, 16, 512, 60000, 1, 1, seed=33, kind="pixels")
    coords, lat, A, S, g = _dev(c["coords"]), _dev(c["lat"]), _dev(c["A"]), _dev(c["S"]), _dev(c["g"])
    plan = lib.Plan(coords)
    feats = lib.latent_forward_planned(plan, lat, c["first"], c["res"], 16, A, S, 1, True)
    gl, gA, gS = lib.latent_backward_planned(plan, g, lat, c["first"], c["res"], 16, A, 1, 1, c["T"], True, True)
    plan2 = lib.Plan(coords).set_sorted_io(True)
    perm = plan2.perm_tensor()
    assert torch.equal(torch.sort(perm)[0], torch.arange(coords.shape[0], device="cuda"))
    feats_s = lib.latent_forward_planned(plan2, lat, c["first"], c["res"], 16, A, S, 1, True)
    assert torch.equal(feats_s, feats[perm])
    gl_s, gA_s, gS_s = lib.latent_backward_planned(plan2, g[perm].contiguous(), lat, c["first"], c["res"], 16, A, 1, 1,
                                                   c["T"], True, True)
    # per tile the fixed-point sums are order independent; nodes shared by neighbouring tiles receive their few
    # float adds in launch order, hence last-bit differences between two launches
    assert rel_err(gl_s.cpu().numpy(), gl.cpu().numpy()) <= 1e-6
    assert rel_err(gA_s.cpu().numpy().sum(0), gA.cpu().numpy().sum(0)) <= 1e-5
    assert rel_err(gS_s.cpu().numpy().sum(0), gS.cpu().numpy().sum(0)) <= 1e-5
    plan.close()
    plan2.close()


def test_drop_in_entry_points_take_the_tiled_path_for_large_2d_batches(lib):
    """wisp._C.ops-compatible entry points: a plain table is the latent grid with identity decoder and no rounding,
    so large 2D batches run on the tiled kernels; values equal the point-parallel kernels', gradients agree."""
    from shacira_b200 import grid_ops
    from shacira_b200._C import ops
    c = _case(2, 16, 16, 16, 512, 70000, 2, 2, seed=12, kind="pixels")
    coords, g = _dev(c["coords"]), _dev(c["g"])
    table = _dev(np.random.default_rng(1).standard_normal((c["T"], 2)).astype(np.float32))
    first = torch.tensor(c["first"], dtype=torch.int32, device="cuda")
    grid_ops.clear_plans()
    before = dict(grid_ops.plan_stats)
    feats = ops.hashgrid_interpolate2d_cuda(coords, table, first, c["res"], 16)
    assert grid_ops.plan_stats["builds"] == before["builds"] + 1                   # the tiled path ran
    want = lib.hashgrid_forward(coords, table, c["first"], c["res"], 16)           # point-parallel kernel
    assert torch.equal(feats, want)
    gt = ops.hashgrid_interpolate2d_backward_cuda(coords, g, table, first, c["res"], 16, 2, False)
    assert grid_ops.plan_stats["hits"] >= before["hits"] + 1                        # the forward's plan, reused
    want_g = lib.hashgrid_backward(coords, g, c["first"], c["res"], 16, 2, c["T"])
    assert rel_err(gt.cpu().numpy(), want_g.cpu().numpy()) <= BWD_TOL
    # autograd through the reference-named op
    t2 = table.clone().requires_grad_(True)
    out = grid_ops.hashgrid2d(coords, c["res"], 16, 0, t2, None, first)
    out.backward(g)
    assert rel_err(t2.grad.cpu().numpy(), want_g.cpu().numpy()) <= BWD_TOL
    # 3D and small batches stay on the point-parallel kernels
    builds = grid_ops.plan_stats["builds"]
    small = ops.hashgrid_interpolate2d_cuda(coords[:1000].contiguous(), table, first, c["res"], 16)
    assert grid_ops.plan_stats["builds"] == builds and torch.equal(small, want[:1000])


def test_host_session_pipelined_steps_match_device_calls(lib):
    """shacira_host_session_*: host buffers in and out, two steps in flight, tables that change between steps: every
    step's features, table gradient and decoder gradients equal the device-tensor calls on the same inputs."""
    c = _case(2, 16, 16, 16, 512, 40000, 1, 1, seed=21, kind="uniform")
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    coords = pin(c["coords"])
    sess = lib.HostSession(coords, c["T"], c["first"], c["res"], 16, 1, 1)
    rng = np.random.default_rng(0)
    K = 5
    lats = [pin(c["lat"] + rng.standard_normal(c["lat"].shape).astype(np.float32) * k) for k in range(K)]
    gouts = [pin(rng.standard_normal(c["g"].shape).astype(np.float32)) for _ in range(K)]
    A, S = pin(c["A"]), pin(c["S"])
    feats = [torch.empty(c["g"].shape, dtype=torch.float32).pin_memory() for _ in range(K)]
    gls = [torch.empty(c["lat"].shape, dtype=torch.float32).pin_memory() for _ in range(K)]
    gAs = [torch.empty((16, 1, 1), dtype=torch.float32).pin_memory() for _ in range(K)]
    gSs = [torch.empty((16, 1), dtype=torch.float32).pin_memory() for _ in range(K)]
    slots = []
    for k in range(K):
        sess.set_table(lats[k], A, S, True)
        slots.append(sess.step(gouts[k], feats[k], gls[k], gAs[k], gSs[k]))
        if k >= 1:
            sess.wait(slots[k - 1])
    sess.wait(slots[-1])
    assert slots == [0, 1, 0, 1, 0]
    dcoords, dA, dS = coords.cuda(), A.cuda(), S.cuda()
    plan = lib.Plan(dcoords)
    for k in range(K):
        f = lib.latent_forward_planned(plan, lats[k].cuda(), c["first"], c["res"], 16, dA, dS, 1, True)
        gl, gA, gS = lib.latent_backward_planned(plan, gouts[k].cuda(), lats[k].cuda(), c["first"], c["res"], 16, dA, 1, 1,
                                                 c["T"], True, True)
        assert torch.equal(feats[k], f.cpu())
        assert rel_err(gls[k].numpy(), gl.cpu().numpy()) <= BWD_TOL   # another plan object: other point order inside a tile, other batches and scales
        # only the SUM over the 16 rows is defined (tiles spread their share over the rows): the rows are O(10^3) and
        # cancel to O(1), so the sum is compared against the magnitude of what was added, in float64
        for got, ref in ((gAs[k], gA), (gSs[k], gS)):
            a, b = got.numpy().astype(np.float64), ref.cpu().numpy().astype(np.float64)
            assert abs(a.sum() - b.sum()) <= 1e-5 * max(np.abs(a).sum(), np.abs(b).sum())
    plan.close()
    sess.close()


@pytest.mark.parametrize("L,bw,n", [(16, 16, 768 * 512), (24, 11, 768 * 512)])
def test_self_scaled_backward_per_level_and_per_tile(lib, L, bw, n):
    """The tiled backward accumulates in fixed point with a step of 2^-19 of each TILE's and LEVEL's own max |gradient|.
    A gate normalised by the whole table's maximum would let a small gradient next to a large one be wrong by far more
    than 1e-4 of itself, so: upstream gradients whose magnitude varies by 10^6 across the image, and the error
    normalised (a) per level and (b) per level AND spatial tile (dense levels: a row is a grid node, a node lies in a
    tile), plus RMS per level. BASELINE cfg2 and the reference's kodak.yaml grid (24 levels, 2^11 rows)."""
    from helpers import level_rel_err, level_rms_err
    c = _case(2, L, bw, 16, 512, n, 1, 1, seed=77, kind="pixels")
    # gradient magnitude: 10^-3 ... 10^3 across the image, smooth in space (whole tiles are small or large)
    mag = 10.0 ** (3.0 * np.sin(2.5 * c["coords"][:, 0]) * np.cos(1.7 * c["coords"][:, 1]))
    c["g"] = (c["g"] * mag[:, None]).astype(np.float32)
    coords, lat, A, g = _dev(c["coords"]), _dev(c["lat"]), _dev(c["A"]), _dev(c["g"])
    plan = lib.Plan(coords)
    G = plan.info()["tiles_per_axis"]
    gl, _, _ = lib.latent_backward_planned(plan, g, None, c["first"], c["res"], bw, A, 1, 1, c["T"], True, False)
    gl = gl.cpu().numpy().astype(np.float64)
    _, want_gl, _, _ = _oracle_fwd_bwd(c)
    sizes = oracle.level_layout(c["res"], bw, 2)[0]
    assert level_rel_err(gl, want_gl, c["first"], sizes) <= BWD_TOL
    assert level_rms_err(gl, want_gl, c["first"], sizes) <= BWD_TOL
    checked = 0
    for l, (f0, rows, res) in enumerate(zip(c["first"], sizes, c["res"])):
        if res * res != rows:          # hashed level: rows are not spatial
            continue
        a = gl[f0:f0 + rows, 0].reshape(res, res)            # row = x + y * res  ->  [y, x]
        b = want_gl[f0:f0 + rows, 0].reshape(res, res)
        # node i along an axis lies at coordinate i / res in [0, 1): tile = floor(i * G / res)
        tile_of = np.minimum((np.arange(res) * G) // res, G - 1)
        for ty in range(G):
            ys = np.nonzero(tile_of == ty)[0]
            for tx in range(G):
                xs = np.nonzero(tile_of == tx)[0]
                if ys.size == 0 or xs.size == 0:
                    continue
                ba = b[np.ix_(ys, xs)]
                m = np.abs(ba).max()
                if m == 0.0:
                    continue
                # a node receives contributions from every tile within one CELL of it, each quantised with that tile's
                # own scale: on a coarse level (cell = G / res tiles wide) that is several tiles. Normalise by the
                # largest gradient over the tiles that can contribute (magnitudes vary by ~3x from one tile to the
                # next here, by 10^6 across the image)
                rad = int(np.ceil(G / res)) + 1
                ny = np.nonzero((tile_of >= ty - rad) & (tile_of <= ty + rad))[0]
                nx = np.nonzero((tile_of >= tx - rad) & (tile_of <= tx + rad))[0]
                m = np.abs(b[np.ix_(ny, nx)]).max()
                err = np.abs(a[np.ix_(ys, xs)] - ba).max() / m
                assert err <= BWD_TOL, "level %d tile (%d, %d): %.3g" % (l, tx, ty, err)
                checked += 1
    assert checked > 100
    plan.close()
