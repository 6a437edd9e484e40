"""Host-side mirror of the reference interface: construction, layouts, names, decoders,
bitstream. CPU only (no kernels are launched)."""
import numpy as np
import pytest
import torch

import oracle
from helpers import affine_from_case, case_from_golden


def _dec_cfg(kind="single", C=1, **kw):
    d = dict(ldecode_enabled=True, ldecode_type=kind, use_sga=False, diff_sampling=True, use_shift=True,
             ldecode_matrix="sq", latent_dim=C, norm="max", norm_every=10, ldec_std=0.1, decay_period=0.9,
             temperature=0.1)
    d.update(kw)
    return d


def _ent_cfg(layers=2):
    return dict(num_prob_layers=layers, entropy_reg=1e-3, entropy_reg_end=1e-4, entropy_reg_sched="cosine",
                noise_freq=1)


def test_geometric_levels_match_survey_cfg2(lib):
    from shacira_b200.grids import LatentGrid
    g = LatentGrid.from_geometric(feature_dim=1, num_lods=16, latent_dim=1, multiscale_type="cat", resolution_dim=2,
                                  feature_std=0.1, codebook_bitwidth=16, min_grid_res=16, max_grid_res=512,
                                  init_grid="uniform", conf_latent_decoder=_dec_cfg(), conf_entropy_reg=_ent_cfg())
    assert g.resolutions == [17, 21, 26, 33, 41, 51, 65, 81, 102, 129, 162, 204, 257, 323, 407, 513]
    assert g.codebook.shape == (374612, 1)
    assert g.codebook_lod_sizes.dtype == torch.int32 and g.codebook_lod_first_idx.dtype == torch.int32
    assert g.codebook_lod_sizes.tolist()[-4:] == [65536] * 4
    assert float(g.codebook.detach().abs().max()) <= float(np.float32(0.1))  # uniform(+-feature_std) in float32


def test_state_dict_names_match_reference(lib):
    """the trainer's optimizer grouping keys on these names."""
    from shacira_b200.grids import LatentGrid
    g = LatentGrid.from_geometric(feature_dim=1, num_lods=4, latent_dim=1, multiscale_type="cat", resolution_dim=2,
                                  codebook_bitwidth=8, min_grid_res=4, max_grid_res=32,
                                  conf_latent_decoder=_dec_cfg(), conf_entropy_reg=_ent_cfg())
    names = set(g.state_dict().keys())
    want = {"codebook", "codebook_lod_sizes", "codebook_lod_first_idx", "latent_dec.div",
            "latent_dec.layers.0.scale", "latent_dec.layers.0.shift"}
    want |= {"prob_model.f%d.%s" % (i, p) for i in (1, 2, 3) for p in "hba"} | {"prob_model.f4.h", "prob_model.f4.b"}
    assert names == want
    assert not g.latent_dec.div.requires_grad
    gh = LatentGrid.from_geometric(feature_dim=4, num_lods=3, latent_dim=2, multiscale_type="cat", resolution_dim=3,
                                   codebook_bitwidth=8, min_grid_res=4, max_grid_res=16,
                                   conf_latent_decoder=_dec_cfg("hierarchical", 2), conf_entropy_reg=_ent_cfg())
    assert "latent_dec.decoders.2.layers.0.scale" in gh.state_dict()
    assert gh.latent_dec.decoders[0].layers[0].scale.shape == (2, 4)


def test_golden_state_loads_into_mirror_and_decodes_identically(lib, golden):
    """Reference parameters dropped into the mirror classes reproduce the reference's decoded table."""
    from shacira_b200.latent_decoders import LatentDecoder
    c = case_from_golden(golden, "nerf_c4f4_dft")
    dec = LatentDecoder(**_dec_cfg("single", 4, ldecode_matrix="dft", feature_dim=4))
    dec.load_state_dict({"div": torch.from_numpy(c["div0"]), "layers.0.scale": torch.from_numpy(c["scale0"]),
                         "layers.0.shift": torch.from_numpy(c["shift0"]), "layers.0.dft": torch.from_numpy(c["dft0"])})
    assert np.array_equal(dec.layers[0].dft.numpy(), c["dft0"])  # same DCT basis as the reference builds
    table = dec(torch.from_numpy(c["codebook"])).detach().numpy()
    assert np.array_equal(table, c["table"])
    A, shift = dec.affine_map()
    A_ref, S_ref = affine_from_case(c)
    assert np.array_equal(A.detach().numpy(), A_ref) and np.array_equal(shift.detach().numpy(), S_ref)


def test_affine_map_equals_forward(lib):
    from shacira_b200.latent_decoders import HierarchicalLatentDecoder, LatentDecoder
    torch.manual_seed(0)
    dec = LatentDecoder(**_dec_cfg("single", 2, feature_dim=4, ldec_std=0.5))
    dec.div.data = torch.tensor([2.0, 0.5])
    w = torch.randn(50, 2) * 5
    A, shift = dec.affine_map()
    assert torch.allclose(torch.round(w) @ A[0] + shift, dec(w), atol=1e-6)
    assert not LatentDecoder(**_dec_cfg(feature_dim=1, final_activation="tanh")).is_affine()
    assert not LatentDecoder(**_dec_cfg(feature_dim=1, num_layers_dec=1)).is_affine()
    h = HierarchicalLatentDecoder(3, [0, 10, 30, 50], _dec_cfg("hierarchical", 2, feature_dim=4))
    A, shift = h.affine_map()
    assert A.shape == (3, 2, 4) and shift.shape == (3, 4)
    out = h(w)
    assert torch.allclose(out[10:30], torch.round(w[10:30]) @ A[1] + shift[1], atol=1e-6)
    assert torch.isfinite(out).all()  # last level decoded (SURVEY Q5 fenced)


def test_straight_through_gradients():
    from shacira_b200.latent_decoders import StraightThrough, StraightThroughFloor
    w = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.2], requires_grad=True)
    q = StraightThrough.apply(w)
    assert q.tolist() == [0.0, 2.0, 2.0, -0.0, -1.0]  # half to even, like torch.round
    q.sum().backward()
    assert w.grad.tolist() == [1.0] * 5
    w.grad = None
    StraightThroughFloor.apply(w).sum().backward()
    assert w.grad.tolist() == [1.0] * 5


def test_sga_quantize_bounds_and_gradient():
    from shacira_b200.latent_decoders import sga_quantize
    torch.manual_seed(1)
    w = (torch.rand(1000, 1) * 10 - 5).requires_grad_(True)
    q = sga_quantize(w, 0.5, diff_sampling=True)
    assert bool(((q >= torch.floor(w) - 1e-4) & (q <= torch.floor(w) + 1 + 1e-4)).all())
    q.sum().backward()
    assert w.grad is not None and torch.isfinite(w.grad).all()


def test_prob_model_matches_oracle_and_packs(lib, golden):
    from oracle import latent_oracle as lo
    from shacira_b200.prob_models import BitEstimator
    torch.manual_seed(2)
    pm = BitEstimator(3, num_layers=4)
    x = torch.randn(20, 3) * 4
    params = {"f%d" % i: (f.h, f.b, f.a) for i, f in enumerate((pm.f1, pm.f2, pm.f3, pm.f4), 1)}
    assert torch.equal(pm(x), lo.bit_estimator(x, params, 4))
    assert torch.allclose(pm(x[:, 1], single_channel=1), pm(x)[:, 1], atol=1e-6)
    packed = pm.packed_params()
    assert packed.shape == (4, 3, 3) and torch.equal(packed[3, 2], torch.zeros(3))
    assert torch.equal(packed[1, 0], pm.f2.h[0])


def test_bitstream_round_trip_and_length(lib):
    from shacira_b200 import bitstream
    torch.manual_seed(3)
    col = torch.round(torch.randn(20000) * 3)
    uniq, counts = torch.unique(col.long(), return_counts=True)
    stream, cdf = bitstream.encode_column(col, uniq, counts)
    back = bitstream.decode_column(stream, cdf, col.numel(), uniq)
    assert torch.equal(back, col.long())
    p = counts / counts.sum()
    entropy_bits = float(-(counts * torch.log2(p)).sum())
    assert entropy_bits <= len(stream) * 8 <= entropy_bits * 1.01 + 64
    # what the reference would hand to torchac: identical symbols and float CDF (oracle restatement)
    from oracle import latent_oracle as lo
    sym, cdf_f, _, _ = lo.symbol_stream(col)
    assert torch.equal(bitstream.dense_ranks(col, uniq), sym)
    assert torch.equal(bitstream.float_cdf(counts), cdf_f)
    # the container's coding model: exact integer arithmetic on the counts, strictly increasing, total 2^16; it does
    # not depend on float summation order (a permuted cumsum of the float CDF may differ in the last bit)
    icdf = bitstream.integer_cdf(counts)
    assert icdf[0] == 0 and icdf[-1] == 65536 and (np.diff(icdf.astype(np.int64)) > 0).all()
    total = int(counts.sum())
    want = [(int(c) * (65536 - len(counts))) // total + k for k, c in enumerate([0] + np.cumsum(counts.numpy()).tolist())]
    assert icdf.tolist() == want
    s2, cdf2 = bitstream.encode_column(col, uniq, counts, exact=True)
    assert np.array_equal(cdf2, icdf) and torch.equal(bitstream.decode_column(s2, icdf, col.numel(), uniq), col.long())
    # degenerate: a single symbol
    one = torch.zeros(100)
    u1, c1 = torch.unique(one.long(), return_counts=True)
    s1, cdf1 = bitstream.encode_column(one, u1, c1)
    assert torch.equal(bitstream.decode_column(s1, cdf1, 100, u1), one.long())


def test_dp_sharding_helpers():
    from shacira_b200 import dp
    units = [dp.shard_units(24, r, 8) for r in range(8)]
    assert sorted(sum(units, [])) == list(range(24)) and all(len(u) == 3 for u in units)
    spans = [dp.split_rays(4096 + 3, r, 4) for r in range(4)]
    assert spans[0][0] == 0 and spans[-1][1] == 4099
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_wisp_C_drop_in_names(lib):
    from shacira_b200 import _C, compat
    for n in ("hashgrid_interpolate_cuda", "hashgrid_interpolate_backward_cuda", "hashgrid_interpolate2d_cuda",
              "hashgrid_interpolate2d_backward_cuda"):
        assert callable(getattr(_C.ops, n))
    import sys
    saved = {k: sys.modules.get(k) for k in ("wisp._C", "wisp._C.ops")}
    try:
        assert compat.install_as_wisp_C() is _C and sys.modules["wisp._C.ops"] is _C.ops
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_odd_feature_dim_raises_like_reference(lib):
    from shacira_b200 import grid_ops
    with pytest.raises(Exception, match="multiple of 2"):
        grid_ops.hashgrid2d(torch.zeros(4, 2), [17], 10, 0, torch.zeros(289, 1), None, torch.zeros(1, dtype=torch.int32))


def test_codec_container_round_trip_and_size_accounting(lib):
    """the fitted model as a real byte stream. Decoding restores round(latents) and every decoder /
    MLP parameter exactly; the file is the reference's BPP formula (image_trainer.py:162-168) plus the histogram and
    header the formula leaves out."""
    import torch.nn as nn
    from shacira_b200 import codec
    from shacira_b200.grids import LatentGrid
    torch.manual_seed(11)
    dec = dict(ldecode_enabled=True, ldecode_type="single", use_sga=False, diff_sampling=True, use_shift=True,
               ldecode_matrix="sq", latent_dim=2, norm="max", norm_every=10, ldec_std=0.1, decay_period=0.9,
               temperature=0.1)
    ent = dict(num_prob_layers=2, entropy_reg=1e-3, entropy_reg_end=1e-4, entropy_reg_sched="cosine", noise_freq=1)
    mk = lambda: LatentGrid.from_geometric(feature_dim=4, num_lods=8, latent_dim=2, multiscale_type="cat",
                                           resolution_dim=2, feature_std=0.1, codebook_bitwidth=12, min_grid_res=16,
                                           max_grid_res=128, init_grid="uniform", conf_latent_decoder=dict(dec),
                                           conf_entropy_reg=dict(ent))
    grid = mk()
    mlp = nn.Sequential(nn.Linear(32, 16), nn.ReLU(), nn.Linear(16, 16), nn.ReLU(), nn.Linear(16, 3))
    with torch.no_grad():
        grid.codebook.copy_(torch.randn_like(grid.codebook) * torch.tensor([3.0, 9.0]))
        grid.latent_dec.div.copy_(torch.tensor([2.5, 7.0]))
    blob = codec.encode_model(grid, mlp, extra_meta={"image": "synthetic"})
    state = codec.decode_model(blob)
    assert torch.equal(state["latents"], torch.round(grid.codebook.detach()).long())
    assert state["header"]["meta"] == {"image": "synthetic"} and state["header"]["num_lods"] == 8
    grid2, mlp2 = mk(), nn.Sequential(nn.Linear(32, 16), nn.ReLU(), nn.Linear(16, 16), nn.ReLU(), nn.Linear(16, 3))
    codec.load_into(state, grid2, mlp2)
    assert torch.equal(grid2.codebook.detach(), torch.round(grid.codebook.detach()))
    for (k1, v1), (k2, v2) in zip(grid.latent_dec.state_dict().items(), grid2.latent_dec.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    for v1, v2 in zip(mlp.state_dict().values(), mlp2.state_dict().values()):
        assert torch.equal(v1, v2)
    # the decoder path only ever sees round(w): the decoded table decodes to the same features
    assert torch.equal(grid2.latent_dec(grid2.codebook), grid.latent_dec(grid.codebook))
    rep = codec.size_report(grid, mlp, blob, pixels=768 * 512)
    assert rep["file_bytes"] == len(blob)
    assert rep["latent_entropy_bits"] <= rep["latent_stream_bits"] <= rep["latent_entropy_bits"] * 1.01 + 128
    # file = formula + (histogram, header, framing), and nothing else
    assert 0 < rep["uncounted_by_reference_bits"] <= rep["histogram_bits"] + 8 * 2048
    # corruption is detected, not silently decoded
    with pytest.raises(ValueError):
        codec.decode_model(b"XXXX" + blob[4:])
    with pytest.raises(ValueError):
        codec.decode_model(blob + b"\0")
    # a flipped bit inside the coded latents: the per-channel CRC (or the decoder itself) refuses the stream
    import json
    import struct
    hlen = struct.unpack_from("<HI", blob, 4)[1]
    pos = 10 + hlen
    lo0, K0 = struct.unpack_from("<iI", blob, pos)
    first_stream = pos + 8 + 4 * K0 + 12
    bad = bytearray(blob)
    bad[first_stream + 40] ^= 0x10
    with pytest.raises((ValueError, lib.ShaciraError)):
        codec.decode_model(bytes(bad))


def test_first_idx_host_cache_survives_address_reuse():
    """The shim caches the host copy of a device / tensor first_idx (the reference passes a tensor on every call). A new
    tensor that the allocator places at a freed tensor's address must not be served the old contents."""
    import torch
    from shacira_b200._C.ops import _host_ints
    for i in range(200):
        t = torch.arange(16, dtype=torch.int32) * (i + 1)
        assert _host_ints(t) == tuple(int(v) for v in t.tolist())
        del t
