"""Fused latent grid (quantise -> gather -> lerp -> decode) and the bit-rate loss on a B200,
against the golden vectors produced by the reference's own Python modules and the CPU oracle."""
import numpy as np
import pytest
import torch

import oracle
from oracle import latent_oracle as lo
from helpers import affine_from_case, case_from_golden, prob_params_from_case, rel_err

pytestmark = pytest.mark.gpu

CASES = ["img_c1f1", "img_c2f4_h", "nerf_c1f4", "nerf_c4f4_dft"]
FWD_TOL, BWD_TOL = 1e-5, 1e-4


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _dec_cfg(c, use_sga=False):
    return dict(ldecode_enabled=True, ldecode_type="hierarchical" if c["hier"] else "single", use_sga=use_sga,
                diff_sampling=True, use_shift=True, ldecode_matrix="dft" if c["dft"] else "sq", latent_dim=c["C"],
                norm="max", norm_every=10, ldec_std=0.1, decay_period=0.9, temperature=0.1)


def _ent_cfg(c):
    return dict(num_prob_layers=c["layers"], entropy_reg=1e-3, entropy_reg_end=1e-4, entropy_reg_sched="cosine",
                noise_freq=2)


def _grid_from_case(c, use_sga=False):
    """Our LatentGrid with the reference's parameters loaded through state_dict names."""
    from shacira_b200.grids import LatentGrid
    grid = LatentGrid.from_resolutions(feature_dim=c["F"], resolutions=c["resolutions"], latent_dim=c["C"],
                                       multiscale_type="cat", resolution_dim=c["dim"], feature_std=0.1,
                                       codebook_bitwidth=c["bw"], init_grid="uniform",
                                       conf_latent_decoder=_dec_cfg(c, use_sga), conf_entropy_reg=_ent_cfg(c))
    sd = grid.state_dict()
    sd["codebook"] = torch.from_numpy(c["codebook"])
    nA = c["L"] if c["hier"] else 1
    for i in range(nA):
        pre = ("latent_dec.decoders.%d." % i) if c["hier"] else "latent_dec."
        sd[pre + "div"] = torch.from_numpy(c["div%d" % i])
        sd[pre + "layers.0.scale"] = torch.from_numpy(c["scale%d" % i])
        sd[pre + "layers.0.shift"] = torch.from_numpy(c["shift%d" % i])
    for fi in range(4):
        sd["prob_model.f%d.h" % (fi + 1)] = torch.from_numpy(c["prob_h%d" % fi])
        sd["prob_model.f%d.b" % (fi + 1)] = torch.from_numpy(c["prob_b%d" % fi])
        if fi < 3:
            sd["prob_model.f%d.a" % (fi + 1)] = torch.from_numpy(c["prob_a%d" % fi])
    grid.load_state_dict(sd)
    assert grid.codebook_lod_first_idx.tolist() == c["first_idx"]
    return grid.cuda()


@pytest.mark.parametrize("name", CASES)
def test_fused_kernel_forward_matches_reference(lib, golden, name):
    c = case_from_golden(golden, name)
    A, S = affine_from_case(c)
    feats, z = lib.latent_forward(_dev(c["coords"]), _dev(c["codebook"]), c["first_idx"], c["resolutions"], c["bw"],
                                  _dev(A), _dev(S), c["F"], True, True)
    assert rel_err(feats.cpu().numpy(), c["feats"]) <= FWD_TOL
    # z is the interpolation of the ROUNDED latents: oracle on rint(codebook)
    zq = oracle.forward(c["coords"], np.rint(c["codebook"]).astype(np.float32), c["first_idx"], c["resolutions"], c["bw"])
    assert rel_err(z.cpu().numpy(), zq) <= FWD_TOL
    # without rounding it is the plain interpolation of the raw latents
    _, z_raw = lib.latent_forward(_dev(c["coords"]), _dev(c["codebook"]), c["first_idx"], c["resolutions"], c["bw"],
                                  _dev(A), None, c["F"], False, True)
    zr = oracle.forward(c["coords"], c["codebook"], c["first_idx"], c["resolutions"], c["bw"])
    assert rel_err(z_raw.cpu().numpy(), zr) <= FWD_TOL


@pytest.mark.parametrize("name", CASES)
def test_latent_grid_module_matches_reference_fwd_bwd(lib, golden, name):
    """The public API a user of the reference calls: LatentGrid.interpolate + autograd."""
    c = case_from_golden(golden, name)
    grid = _grid_from_case(c)
    feats = grid.interpolate(_dev(c["coords"]), 0)
    assert feats.shape == c["feats"].shape
    assert rel_err(feats.detach().cpu().numpy(), c["feats"]) <= FWD_TOL
    feats.backward(_dev(c["grad_out"]))
    assert rel_err(grid.codebook.grad.cpu().numpy(), c["grad_codebook"]) <= BWD_TOL
    decs = grid.latent_dec.decoders if c["hier"] else [grid.latent_dec]
    for i, d in enumerate(decs):
        assert rel_err(d.layers[0].scale.grad.cpu().numpy(), c["grad_scale%d" % i]) <= BWD_TOL
        assert rel_err(d.layers[0].shift.grad.cpu().numpy(), c["grad_shift%d" % i]) <= BWD_TOL


@pytest.mark.parametrize("name", CASES)
def test_fused_equals_unfused_composition_on_gpu(lib, golden, name):
    """Table-side decode (PyTorch, GPU) + plain kernel == fused kernel."""
    from shacira_b200 import grid_ops
    c = case_from_golden(golden, name)
    grid = _grid_from_case(c)
    coords = _dev(c["coords"])
    fused = grid.interpolate(coords, 0)
    table = grid.latent_dec(grid.codebook)
    plain = grid_ops.hashgrid_any(coords, table, c["first_idx"], c["resolutions"], c["bw"])
    assert rel_err(fused.detach().cpu().numpy(), plain.detach().cpu().numpy()) <= FWD_TOL


def test_sga_mode_runs_and_differentiates(lib, golden):
    c = case_from_golden(golden, "img_c1f1")
    grid = _grid_from_case(c, use_sga=True)
    grid.latent_dec.temperature = 0.5
    torch.manual_seed(0)
    feats = grid.interpolate(_dev(c["coords"]), 0)
    assert feats.shape == c["feats"].shape and torch.isfinite(feats).all()
    # SGA output lies between the decodes of floor and ceil; here just the pipeline + gradient flow
    feats.backward(_dev(c["grad_out"]))
    assert grid.codebook.grad is not None and torch.isfinite(grid.codebook.grad).all()
    assert float(grid.codebook.grad.abs().sum()) > 0
    assert grid.latent_dec.layers[0].scale.grad is not None


def test_non_affine_decoder_falls_back_to_table_side_decode_on_gpu(lib, golden):
    from shacira_b200.grids import LatentGrid
    c = case_from_golden(golden, "nerf_c1f4")
    cfg = _dec_cfg(c)
    cfg["final_activation"] = "tanh"
    grid = LatentGrid.from_resolutions(feature_dim=4, resolutions=c["resolutions"], latent_dim=1, multiscale_type="cat",
                                       resolution_dim=3, feature_std=0.1, codebook_bitwidth=c["bw"],
                                       conf_latent_decoder=cfg, conf_entropy_reg=_ent_cfg(c)).cuda()
    with torch.no_grad():
        grid.codebook.copy_(_dev(c["codebook"]))
    feats = grid.interpolate(_dev(c["coords"]), 0)
    table = torch.tanh(torch.round(grid.codebook) @ grid.latent_dec.layers[0].scale + grid.latent_dec.layers[0].shift)
    want = oracle.forward(c["coords"], table.detach().cpu().numpy(), c["first_idx"], c["resolutions"], c["bw"])
    assert rel_err(feats.detach().cpu().numpy(), want) <= FWD_TOL
    feats.sum().backward()
    assert grid.codebook.grad is not None


def test_batched_sample_shape_and_sum_aggregation(lib, golden):
    c = case_from_golden(golden, "nerf_c1f4")
    grid = _grid_from_case(c)
    coords = _dev(c["coords"][:1200]).reshape(30, 40, 3)
    out = grid.interpolate(coords, 0)
    assert out.shape == (30, 40, c["L"] * c["F"])
    assert rel_err(out.reshape(1200, -1).detach().cpu().numpy(), c["feats"][:1200]) <= FWD_TOL
    grid.multiscale_type = "sum"
    s = grid.interpolate(coords, 0)
    assert s.shape == (30, 40, c["F"])
    assert torch.allclose(s, out.reshape(30, 40, c["L"], c["F"]).sum(-2), atol=1e-5)


# ---- bit-rate loss -------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["train", "val"])
def test_entropy_kernel_matches_reference(lib, golden, name, mode):
    c = case_from_golden(golden, name)
    packed, _ = prob_params_from_case(c)
    noise = None if mode == "val" else _dev(c["noise"])
    bits, gl, gp = lib.entropy_bits(_dev(c["codebook"]), noise, _dev(packed), c["layers"], c["first_idx"])
    bits = bits.cpu().numpy()
    ref_total = c["ent_%s_total" % mode][0]
    assert abs(bits[0] - ref_total) <= 1e-5 * ref_total           # bpp gate is 1 %; this is 1e-5
    assert abs(bits[1:].sum() - bits[0]) <= 1e-6 * bits[0]        # per-level reduction adds up
    assert rel_err(gl.cpu().numpy(), c["ent_%s_grad_codebook" % mode]) <= BWD_TOL or mode == "val"
    if mode == "val":
        assert not bool(gl.any())                                  # round() passes no gradient
    gp = gp.cpu().numpy()
    m = min(c["layers"], 4) - 1
    used = list(range(m)) + [3]
    for fi in used:
        for pi, pn in enumerate("hba"):
            key = "ent_%s_grad_%s%d" % (mode, pn, fi)
            if key in c:
                assert rel_err(gp[fi, pi], c[key].reshape(-1)) <= BWD_TOL, key
    for fi in set(range(4)) - set(used):
        assert not gp[fi].any()                                    # unused layers get no gradient


@pytest.mark.parametrize("C,L", [(1, 16), (2, 6), (4, 0)])
def test_entropy_validation_mode_large_table(lib, C, L):
    """Validation mode (x = round(w): what the NeRF trainer evaluates every step) at a NeRF-size table with a
    sprinkle of huge integers, against the torch restatement of ent_loss (oracle); called twice on the same scratch
    (it must come back clean); the latents' gradient is zero and can be skipped."""
    torch.manual_seed(5 + C)
    T = 600000
    w = torch.randn(T, C) * 9.0
    w[::1000] *= 40.0                                   # a sprinkle of integers far outside [-128, 127]
    w[5, 0], w[6, 0] = -128.0, 127.4                      # the histogram's edges
    first = [int(v) for v in torch.linspace(0, T, L + 1)[:-1]] if L else None
    packed = (torch.randn(4, 3, C) * 0.3).numpy()
    params = {"f%d" % (i + 1): (torch.from_numpy(packed[i, 0:1]), torch.from_numpy(packed[i, 1:2]),
                                torch.from_numpy(packed[i, 2:3]) if i < 3 else None) for i in range(4)}
    wd = w.cuda()
    for layers in (1, 3):
        want = lo.ent_loss(w, None, params, layers, is_val=True)[1].item()
        for _ in range(2):
            bits, gl, gp = lib.entropy_bits(wd, None, torch.from_numpy(packed).cuda(), layers, first, want_grads=True,
                                            want_latent_grads=False)
            assert gl is None
            assert abs(float(bits[0]) - want) <= 1e-5 * want
            if L:
                assert abs(float(bits[1:].sum()) - float(bits[0])) <= 1e-6 * float(bits[0])
        # parameter gradients: autograd through the restatement
        ps = [torch.from_numpy(packed[i].copy()).requires_grad_(True) for i in range(4)]
        pd = {"f%d" % (i + 1): (ps[i][0:1], ps[i][1:2], ps[i][2:3] if i < 3 else None) for i in range(4)}
        lo.ent_loss(w, None, pd, layers, is_val=True)[1].backward()
        got = gp.cpu().numpy()
        for i in list(range(layers - 1)) + [3]:
            g = ps[i].grad.numpy()
            assert rel_err(got[i, :2], g[:2]) <= BWD_TOL
            if i < 3:
                assert rel_err(got[i, 2], g[2]) <= BWD_TOL


@pytest.mark.parametrize("name", CASES)
def test_ent_loss_api_and_autograd(lib, golden, name):
    c = case_from_golden(golden, name)
    grid = _grid_from_case(c)
    grid.noise = _dev(c["noise"])
    avg, tot = grid.ent_loss(1, is_val=False)        # noise_freq=2, odd idx: uses grid.noise like the reference
    ref_total, ref_avg = c["ent_train_total"]
    assert abs(tot.item() - ref_total) <= 1e-5 * ref_total and abs(avg.item() - ref_avg) <= 1e-5 * ref_avg
    (0.5 * avg).backward()
    scale = 0.5 / c["T"]
    assert rel_err(grid.codebook.grad.cpu().numpy(), c["ent_train_grad_codebook"] * scale) <= BWD_TOL
    assert rel_err(grid.prob_model.f4.h.grad.cpu().numpy(), c["ent_train_grad_h3"] * scale) <= BWD_TOL
    # fresh device-side noise every call when noise_freq == 1
    grid.noise_freq = 1
    a1 = grid.ent_loss(0)[1].item()
    a2 = grid.ent_loss(1)[1].item()
    assert a1 != a2 and abs(a1 - a2) < 0.05 * a1


def test_entropy_large_table_against_oracle(lib):
    """BASELINE cfg2 table (374 612 rows): kernel vs the torch-CPU oracle."""
    torch.manual_seed(5)
    T = 374612
    w = torch.randn(T, 1) * 3
    noise = torch.rand(T, 1) - 0.5
    packed = torch.randn(4, 3, 1) * 0.3
    params = {"f%d" % (i + 1): (packed[i, 0:1], packed[i, 1:2], packed[i, 2:3] if i < 3 else None) for i in range(4)}
    _, want = lo.ent_loss(w, noise, params, 2)
    res = oracle.geometric_resolutions(16, 512, 16)
    _, first, _ = oracle.level_layout(res, 16, 2)
    bits, _, _ = lib.entropy_bits(w.cuda(), noise.cuda(), packed.cuda(), 2, first)
    assert abs(bits[0].item() - want.item()) <= 2e-5 * want.item()


# ---- storage size ----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_size_matches_reference(lib, golden, name):
    c = case_from_golden(golden, name)
    grid = _grid_from_case(c)
    ld, cb = grid.size(use_torchac=False)
    assert ld == c["size"][0]
    assert abs(cb - c["size"][1]) <= 1e-6 * c["size"][1]
    ld2, cb2 = grid.size(use_torchac=False, use_prob_model=True)
    assert abs(cb2 - c["size"][3]) <= 1e-4 * c["size"][3]
    # histogram == torch.unique on the rounded column, bit-exact integers
    for ch, (vals, counts) in enumerate(grid.symbol_statistics()):
        u, n = torch.unique(torch.round(torch.from_numpy(c["codebook"][:, ch])).long(), return_counts=True)
        assert torch.equal(vals.cpu(), u) and torch.equal(counts.cpu(), n)
    # the coded stream: length within 1 % (+ a few bytes) of the empirical entropy, decodable
    _, coded = grid.size(use_torchac=True)
    assert cb <= coded <= cb * 1.01 + 64 * c["C"]


def test_quantized_symbols_bit_exact(lib, golden):
    c = case_from_golden(golden, "img_c2f4_h")
    sym, mm = lib.quantize_symbols(_dev(c["codebook"]))
    q = np.rint(c["codebook"]).astype(np.int16)
    assert np.array_equal(sym.cpu().numpy(), q)
    assert mm.cpu().tolist() == [[int(q[:, ch].min()), int(q[:, ch].max())] for ch in range(2)]


# ---- host-buffer entry -----------------------------------------------------------------------------
def test_host_buffer_step_equals_device_path(lib, golden):
    c = case_from_golden(golden, "img_c1f1")
    A, S = affine_from_case(c)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    feats, gl = lib.latent_step_host(pin(c["coords"]), pin(c["codebook"]), c["first_idx"], c["resolutions"], c["bw"],
                                     pin(A), pin(S), pin(c["grad_out"]))
    assert rel_err(feats.numpy(), c["feats"]) <= FWD_TOL
    assert rel_err(gl.numpy(), c["grad_codebook"]) <= BWD_TOL
    # second call reuses the cached device scratch
    feats2, _ = lib.latent_step_host(pin(c["coords"]), pin(c["codebook"]), c["first_idx"], c["resolutions"], c["bw"],
                                     pin(A), pin(S), pin(c["grad_out"]))
    assert torch.equal(feats, feats2)


def test_launch_counter_counts_kernels(lib, golden):
    c = case_from_golden(golden, "img_c1f1")
    before = lib.launch_count()
    lib.hashgrid_forward(_dev(c["coords"]), torch.zeros((c["T"], 2), device="cuda"), c["first_idx"], c["resolutions"], c["bw"])
    assert lib.launch_count() == before + 1   # all levels in ONE launch (the reference: one per level)


def test_in_kernel_noise_matches_buffer_noise_statistically(lib):
    """shacira_entropy_bits_rng: noise from a counter-based hash inside the kernel. Same seed and step -> identical
    result; the step advances by itself; and the estimate agrees with torch-drawn U(-0.5, 0.5) noise to the sampling
    error of 4e5 entries (bits are a mean of i.i.d. terms), as do the parameter gradients."""
    torch.manual_seed(0)
    T, C = 400000, 1
    lat = (torch.randn(T, C, device="cuda") * 3.0)
    params = torch.zeros(4, 3, C, device="cuda")
    params[0, 0], params[0, 1], params[0, 2] = 0.3, 0.1, 0.2
    params[3, 0], params[3, 1] = -0.5, 0.05
    step = torch.zeros((), dtype=torch.int64, device="cuda")
    b0, g0, p0 = lib.entropy_bits_rng(lat, 7, step, params, 2)
    assert int(step) == 1
    b1, g1, p1 = lib.entropy_bits_rng(lat, 7, step, params, 2)
    assert int(step) == 2 and float(b0[0]) != float(b1[0])            # fresh noise every call
    step.zero_()
    b0r, g0r, _ = lib.entropy_bits_rng(lat, 7, step, params, 2)
    assert float(b0r[0]) == float(b0[0]) and torch.equal(g0r, g0)      # counter-based: reproducible
    b_other, _, _ = lib.entropy_bits_rng(lat, 8, torch.zeros((), dtype=torch.int64, device="cuda"), params, 2)
    assert float(b_other[0]) != float(b0[0])
    refs = []
    for _ in range(4):
        noise = torch.rand(T, C, device="cuda") - 0.5
        rb, rg, rp = lib.entropy_bits(lat, noise, params, 2)
        refs.append((float(rb[0]), rp))
    mean_ref = sum(r[0] for r in refs) / len(refs)
    for b in (b0, b1):
        assert abs(float(b[0]) - mean_ref) <= 3e-3 * mean_ref
    assert torch.isfinite(g0).all() and float(g0.abs().max()) > 0
    ref_p = sum(r[1] for r in refs) / len(refs)
    assert float((p0 - ref_p).abs().max()) <= 0.02 * float(ref_p.abs().max())
