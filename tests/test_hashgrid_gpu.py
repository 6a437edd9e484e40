"""Parity of the plain hash-grid kernels (C-ABI) with the CPU oracle on a B200."""
import numpy as np
import pytest
import torch

import oracle
from helpers import make_case, rel_err

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-5   # north_star: forward features within 1e-5 relative (fp32)
BWD_TOL = 1e-4   # north_star: gradients within 1e-4 relative


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


CORNER_CASES = [
    # dim, L, bw, rmin, rmax, n, coord kind
    (2, 16, 14, 16, 512, 1 << 14, "uniform"),     # BASELINE cfg1 grid
    (2, 16, 16, 16, 512, 1 << 14, "pixels"),      # BASELINE cfg2 grid, pixel-centre style coords
    (2, 16, 16, 16, 512, 1 << 14, "arbitrary"),
    (2, 8, 19, 16, 700, 1 << 12, "uniform"),      # dense levels with res >= 257
    (3, 16, 19, 16, 2048, 1 << 14, "arbitrary"),  # BASELINE cfg4 grid, full-mantissa coords
    (3, 16, 14, 16, 512, 1 << 12, "uniform"),
    (3, 24, 19, 16, 512, 1 << 12, "arbitrary"),   # the reference's own lego yaml: 24 levels
]


@pytest.mark.parametrize("dim,L,bw,rmin,rmax,n,kind", CORNER_CASES)
def test_corner_indices_and_weights_bit_exact(lib, dim, L, bw, rmin, rmax, n, kind):
    c = make_case(dim, L, bw, rmin, rmax, n, 2, seed=11, coord_kind=kind)
    coords = c["coords"].copy()
    # edge coordinates: domain corners, centre, just inside, outside the domain (clamped)
    edge = np.array([[-1.0] * dim, [1.0] * dim, [0.0] * dim, [0.9999999] * dim, [-0.9999999] * dim, [1.25] * dim,
                     [-3.0] * dim, [np.nextafter(np.float32(1), np.float32(0))] * dim], dtype=np.float32)
    coords[:len(edge)] = edge
    idx_o, w_o = oracle.corners(coords, c["resolutions"], bw)
    idx_g, w_g = lib.hashgrid_corners(_dev(coords), c["resolutions"], bw)
    idx_g, w_g = idx_g.cpu().numpy(), w_g.cpu().numpy()
    assert np.array_equal(w_g.view(np.uint32), w_o.view(np.uint32)), "weights differ in their bits"
    live = w_o != 0  # zero-weight corners past a dense level are kept in range on the GPU (Q4 fence)
    assert np.array_equal(idx_g[live], idx_o[live])
    sizes = np.array(c["sizes"])[None, :, None]
    assert (idx_g >= 0).all() and (idx_g < sizes).all()
    # the redirected ones exist only where the oracle is out of the level
    assert ((idx_g != idx_o) <= (idx_o >= sizes)).all()


PLAIN_CASES = [
    (2, 16, 14, 16, 512, 4097, 2), (2, 16, 16, 16, 512, 1000, 1), (2, 16, 16, 16, 512, 777, 4),
    (2, 5, 10, 4, 64, 31, 8), (3, 16, 19, 16, 2048, 3001, 2), (3, 16, 19, 16, 2048, 1025, 4),
    (3, 6, 12, 8, 64, 257, 1), (3, 3, 9, 4, 20, 1, 8), (2, 1, 12, 33, 33, 100, 2), (2, 24, 11, 16, 512, 513, 2),
]


@pytest.mark.parametrize("dim,L,bw,rmin,rmax,n,F", PLAIN_CASES)
def test_plain_forward_matches_oracle(lib, dim, L, bw, rmin, rmax, n, F):
    c = make_case(dim, L, bw, rmin, rmax, n, F, seed=dim * 100 + F, coord_kind="arbitrary")
    want = oracle.forward(c["coords"], c["table"], c["first_idx"], c["resolutions"], bw)
    got = lib.hashgrid_forward(_dev(c["coords"]), _dev(c["table"]), c["first_idx"], c["resolutions"], bw).cpu().numpy()
    assert got.shape == want.shape
    assert rel_err(got, want) <= FWD_TOL


@pytest.mark.parametrize("dim,L,bw,rmin,rmax,n,F", PLAIN_CASES[:6])
def test_plain_forward_bit_exact(lib, dim, L, bw, rmin, rmax, n, F):
    c = make_case(dim, L, bw, rmin, rmax, n, F, seed=dim * 100 + F, coord_kind="arbitrary")
    want = oracle.forward(c["coords"], c["table"], c["first_idx"], c["resolutions"], bw)
    got = lib.hashgrid_forward(_dev(c["coords"]), _dev(c["table"]), c["first_idx"], c["resolutions"], bw).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("dim,L,bw,rmin,rmax,n,F", PLAIN_CASES)
def test_plain_backward_matches_oracle(lib, dim, L, bw, rmin, rmax, n, F):
    c = make_case(dim, L, bw, rmin, rmax, n, F, seed=dim * 100 + F + 7, coord_kind="uniform")
    want = oracle.backward(c["coords"], c["grad_out"], c["T"], c["first_idx"], c["resolutions"], bw, F)
    got = lib.hashgrid_backward(_dev(c["coords"]), _dev(c["grad_out"]), c["first_idx"], c["resolutions"], bw, F,
                                c["T"]).cpu().numpy()
    assert rel_err(got, want) <= BWD_TOL


def test_empty_input(lib):
    c = make_case(2, 4, 10, 8, 64, 16, 2, seed=0)
    coords = torch.zeros((0, 2), device="cuda")
    feats = lib.hashgrid_forward(coords, _dev(c["table"]), c["first_idx"], c["resolutions"], 10)
    assert feats.shape == (0, 8)
    g = lib.hashgrid_backward(coords, torch.zeros((0, 8), device="cuda"), c["first_idx"], c["resolutions"], 10, 2, c["T"])
    assert g.shape == (c["T"], 2) and not bool(g.any())


def test_backward_accumulates_into_given_buffer(lib):
    c = make_case(2, 4, 10, 8, 64, 500, 2, seed=1)
    args = (_dev(c["coords"]), _dev(c["grad_out"]), c["first_idx"], c["resolutions"], 10, 2, c["T"])
    g1 = lib.hashgrid_backward(*args)
    buf = torch.ones((c["T"], 2), device="cuda")
    g2 = lib.hashgrid_backward(*args, out=buf)
    assert torch.allclose(g2, g1 + 1, rtol=1e-5, atol=1e-5)


def test_unsupported_and_invalid_arguments_raise(lib):
    c = make_case(2, 2, 10, 8, 16, 10, 2, seed=2)
    with pytest.raises(lib.ShaciraError):
        lib.hashgrid_forward(_dev(c["coords"]), torch.zeros((c["T"], 3), device="cuda"), c["first_idx"], c["resolutions"], 10)
    with pytest.raises(lib.ShaciraError):
        lib.hashgrid_forward(_dev(c["coords"]).double(), _dev(c["table"]), c["first_idx"], c["resolutions"], 10)
    with pytest.raises(lib.ShaciraError):
        lib.hashgrid_forward(torch.zeros((4, 4), device="cuda"), _dev(c["table"]), c["first_idx"], c["resolutions"], 10)


# ---- size-independent properties at BASELINE.json's full sizes ---------------------------------
def _cfg2(n=768 * 512):
    res = oracle.geometric_resolutions(16, 512, 16)
    sizes, first, T = oracle.level_layout(res, 16, 2)
    H, W = 512, 768
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    coords = torch.stack([(ys.reshape(-1) / H - 0.5) * 2, (xs.reshape(-1) / W - 0.5) * 2], 1).float()
    g = torch.Generator().manual_seed(0)
    coords = coords[torch.randperm(coords.shape[0], generator=g)][:n].contiguous().cuda()
    return res, first, T, coords


def test_full_size_partition_of_unity_linearity_adjoint(lib):
    res, first, T, coords = _cfg2()
    n = coords.shape[0]
    ones = lib.hashgrid_forward(coords, torch.ones((T, 2), device="cuda"), first, res, 16)
    assert float((ones - 1).abs().max()) <= 4e-7          # weights sum to 1
    torch.manual_seed(0)
    t1, t2 = torch.randn((T, 2), device="cuda"), torch.randn((T, 2), device="cuda")
    f1 = lib.hashgrid_forward(coords, t1, first, res, 16)
    f2 = lib.hashgrid_forward(coords, t2, first, res, 16)
    f12 = lib.hashgrid_forward(coords, 2 * t1 - t2, first, res, 16)
    assert rel_err((2 * f1 - f2).cpu().numpy(), f12.cpu().numpy()) <= 1e-5  # linear in the table
    g = torch.randn((n, 32), device="cuda")
    gt = lib.hashgrid_backward(coords, g, first, res, 16, 2, T)
    lhs = float((f1.double() * g.double()).sum())
    rhs = float((gt.double() * t1.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), float(n) ** 0.5 * 32)  # backward is the adjoint
    # a checksum of checksums: sum of the gradient equals the sum of the upstream gradient per level
    per_level = gt.double().sum(1)
    for l in range(16):
        a = first[l]
        b = first[l + 1] if l + 1 < 16 else T
        want = float(g[:, 2 * l:2 * l + 2].double().sum())
        assert abs(float(per_level[a:b].sum()) - want) <= 1e-4 * max(abs(want), 1e3)


def test_full_size_3d_sample_against_oracle(lib):
    """BASELINE cfg4 shape (2^19 samples, 16 levels 16->2048, 2^19 rows): oracle-checked on a sample."""
    res = oracle.geometric_resolutions(16, 2048, 16)
    sizes, first, T = oracle.level_layout(res, 19, 3)
    torch.manual_seed(1)
    coords = (torch.rand((1 << 19, 3), device="cuda") * 2 - 1)
    table = torch.randn((T, 2), device="cuda")
    feats = lib.hashgrid_forward(coords, table, first, res, 19)
    pick = torch.randperm(1 << 19, device="cuda")[:4096]
    want = oracle.forward(coords[pick].cpu().numpy(), table.cpu().numpy(), first, res, 19)
    assert rel_err(feats[pick].cpu().numpy(), want) <= FWD_TOL


# ---- the reference-facing Python API ---------------------------------------------------------------
def test_wisp_C_ops_drop_in_signatures(lib):
    from shacira_b200._C import ops
    c = make_case(3, 4, 12, 8, 40, 300, 2, seed=5)
    first_dev = torch.tensor(c["first_idx"], dtype=torch.int32, device="cuda")  # the reference passes a device tensor
    coords, table, gout = _dev(c["coords"]), _dev(c["table"]), _dev(c["grad_out"])
    feats = ops.hashgrid_interpolate_cuda(coords, table, first_dev, c["resolutions"], 12)
    want = oracle.forward(c["coords"], c["table"], c["first_idx"], c["resolutions"], 12)
    assert rel_err(feats.cpu().numpy(), want) <= FWD_TOL
    g = ops.hashgrid_interpolate_backward_cuda(coords, gout, table, first_dev, c["resolutions"], 12, 2, False)
    wantg = oracle.backward(c["coords"], c["grad_out"], c["T"], c["first_idx"], c["resolutions"], 12, 2)
    assert g.shape == table.shape and rel_err(g.cpu().numpy(), wantg) <= BWD_TOL
    c2 = make_case(2, 4, 12, 8, 40, 300, 2, seed=6)
    first_dev = torch.tensor(c2["first_idx"], dtype=torch.int32, device="cuda")
    f2 = ops.hashgrid_interpolate2d_cuda(_dev(c2["coords"]), _dev(c2["table"]), first_dev, c2["resolutions"], 12)
    assert rel_err(f2.cpu().numpy(), oracle.forward(c2["coords"], c2["table"], c2["first_idx"], c2["resolutions"], 12)) <= FWD_TOL


def test_grid_ops_autograd(lib):
    from shacira_b200 import grid_ops
    c = make_case(2, 6, 12, 8, 100, 2000, 2, seed=8)
    first_dev = torch.tensor(c["first_idx"], dtype=torch.int32, device="cuda")
    table = _dev(c["table"]).requires_grad_(True)
    feats = grid_ops.hashgrid2d(_dev(c["coords"]), c["resolutions"], 12, 0, table, None, first_dev)
    assert feats.shape == (2000, 12)
    feats.backward(_dev(c["grad_out"]))
    want = oracle.backward(c["coords"], c["grad_out"], c["T"], c["first_idx"], c["resolutions"], 12, 2)
    assert rel_err(table.grad.cpu().numpy(), want) <= BWD_TOL


def test_hashgrid_module(lib):
    from shacira_b200.grids import HashGrid
    torch.manual_seed(0)
    grid = HashGrid.from_geometric(feature_dim=2, num_lods=8, multiscale_type="cat", resolution_dim=3, feature_std=0.5,
                                   codebook_bitwidth=12, min_grid_res=8, max_grid_res=128).cuda()
    coords = torch.rand((5, 40, 3), device="cuda") * 2 - 1
    out = grid.interpolate(coords, 0)
    assert out.shape == (5, 40, 16)
    want = oracle.forward(coords.reshape(-1, 3).cpu().numpy(), grid.codebook.detach().cpu().numpy(),
                          grid.codebook_lod_first_idx.tolist(), grid.resolutions, 12)
    assert rel_err(out.reshape(-1, 16).detach().cpu().numpy(), want) <= FWD_TOL
    grid.multiscale_type = "sum"
    assert grid.interpolate(coords, 0).shape == (5, 40, 2)
    out.sum().backward()
    assert grid.codebook.grad is not None and grid.size()[1] == grid.codebook.numel() * 32


# ---- large batches: the coarse dense levels are accumulated in shared memory (coarse_kernels.cuh) ---------
COARSE_CASES = [
    # dim, L, bw, rmin, rmax, n, F
    (3, 16, 19, 16, 2048, 1 << 17, 2),   # BASELINE cfg4 grid: 5 dense levels, levels 3 and 4 cut into slabs
    (3, 8, 16, 16, 128, 1 << 16, 4),
    (3, 4, 22, 16, 40, 70001, 1),        # every level dense, ragged n
    (2, 16, 16, 16, 512, 1 << 17, 2),    # 2D point-parallel path: 12 dense levels, slabs along y
    (2, 8, 19, 16, 700, 1 << 16, 2),     # dense levels with res >= 257 (zero-weight corner past the level)
]


@pytest.mark.parametrize("dim,L,bw,rmin,rmax,n,F", COARSE_CASES)
def test_plain_backward_large_batch_matches_oracle(lib, dim, L, bw, rmin, rmax, n, F):
    c = make_case(dim, L, bw, rmin, rmax, n, F, seed=dim * 1000 + F + 3, coord_kind="uniform")
    want = oracle.backward(c["coords"], c["grad_out"], c["T"], c["first_idx"], c["resolutions"], bw, F)
    got = lib.hashgrid_backward(_dev(c["coords"]), _dev(c["grad_out"]), c["first_idx"], c["resolutions"], bw, F,
                                c["T"]).cpu().numpy()
    assert rel_err(got, want) <= BWD_TOL
    # per level too: a coarse level must not hide behind the largest gradient of the whole table
    bounds = list(c["first_idx"]) + [c["T"]]
    for l in range(L):
        a, b = bounds[l], bounds[l + 1]
        assert rel_err(got[a:b], want[a:b]) <= BWD_TOL, "level %d" % l


def test_latent_backward_large_batch_3d_matches_oracle(lib):
    """NeRF shape (C=1 -> F=4, shared affine decoder): grad_latents = scatter of A^T g, against the oracle."""
    dim, L, bw, C, F, n = 3, 16, 19, 1, 4, 1 << 17
    c = make_case(dim, L, bw, 16, 2048, n, F, seed=77, coord_kind="arbitrary")
    rng = np.random.default_rng(5)
    A = rng.standard_normal((1, C, F)).astype(np.float32)
    gz = (c["grad_out"].reshape(n, L, F) @ A[0].T).reshape(n, L * C).astype(np.float32)   # [n, L*C]
    want = oracle.backward(c["coords"], gz, c["T"], c["first_idx"], c["resolutions"], bw, C)
    z = torch.zeros((n, L * C), device="cuda")
    gl, gA, gS = lib.latent_backward(_dev(c["coords"]), _dev(c["grad_out"]), z, c["first_idx"], c["resolutions"], bw,
                                     _dev(A), C, F, c["T"], True)
    assert rel_err(gl.cpu().numpy(), want) <= BWD_TOL
    want_gS = c["grad_out"].reshape(n, L, F).astype(np.float64).sum(0)
    assert rel_err(gS.cpu().numpy(), want_gS) <= BWD_TOL


def test_level_chunked_backward_equals_one_launch(lib):
    """shacira_latent_backward_levels: the backward as a few level chunks (what lets a data-parallel caller overlap
    the all-reduce of one chunk's rows with the next chunk's compute) fills the same buffers as one launch."""
    dim, L, bw, C, F, n = 3, 16, 19, 1, 4, 1 << 17
    c = make_case(dim, L, bw, 16, 2048, n, F, seed=5, coord_kind="uniform")
    rng = np.random.default_rng(6)
    A = rng.standard_normal((1, C, F)).astype(np.float32)
    z = torch.from_numpy(rng.standard_normal((n, L * C)).astype(np.float32)).cuda()
    args = (_dev(c["coords"]), _dev(c["grad_out"]), z, c["first_idx"], c["resolutions"], bw, _dev(A), C, F, c["T"], True)
    gl, gA, gS = lib.latent_backward(*args)
    chunks = [0x000F, 0x03F0, 0xFC00]
    gl2, gA2, gS2 = lib.latent_backward(*args, level_chunks=chunks)
    assert rel_err(gl2.cpu().numpy(), gl.cpu().numpy()) <= 1e-5
    assert rel_err(gA2.cpu().numpy(), gA.cpu().numpy()) <= 1e-5 and rel_err(gS2.cpu().numpy(), gS.cpu().numpy()) <= 1e-5
    # one chunk alone leaves every other level's rows at zero
    gl3, _, gS3 = lib.latent_backward(*args, level_chunks=[0x0030])
    first = list(c["first_idx"]) + [c["T"]]
    assert not bool(gl3[:first[4]].any()) and not bool(gl3[first[6]:].any()) and bool(gl3[first[4]:first[6]].any())
    assert not bool(gS3[:4].any()) and not bool(gS3[6:].any())
