"""Generate tests/golden/sga_ref.npz: the REFERENCE's SGA quantiser (LatentDecoder.forward with use_sga=True,
wisp/models/latent_decoders/basic_latent_decoder.py:183-191, torch's RelaxedOneHotCategorical) on seeded inputs, with the
uniform draws it consumed recorded next to its outputs so that a device kernel can be fed the same noise.

Run in the build container only (`python tests/golden/make_golden_sga.py`); /root/reference is absent on the GPU box.
How the draws are captured: ExpRelaxedCategorical.rsample calls torch.rand(logits.shape) exactly once per forward; the
same seed + the same shape gives the same tensor, so the script draws it first, re-seeds, and runs the reference.
Decoder-level cases use an identity decode (scale = I, div = 1, no shift): the reference's output IS w_hat, and the
gradient of sum(w_out * gout) w.r.t. the latents is gout * d w_hat / d w. The grid-level case runs
LatentGrid.interpolate (table-side SGA -> decode -> interpolation through the C oracle) and its autograd."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import ROOT, dec_cfg, ent_cfg, import_reference  # noqa: E402


def main():
    lg, hg = import_reference()
    from wisp.models.latent_decoders import LatentDecoder
    out = {}
    # ---- decoder level -----------------------------------------------------------------------------------------
    specs = [("t1.0_diff", 3000, 1, 1.0, True), ("t0.37_diff_c2", 2000, 2, 0.37, True), ("t0.1_diff", 3000, 1, 0.1, True),
             ("t0.5_nodiff", 1500, 1, 0.5, False), ("t0.05_diff_c4", 800, 4, 0.05, True)]
    for si, (name, T, C, tau, diff) in enumerate(specs):
        cfg = dec_cfg("single", C, use_shift=False)
        cfg.update(use_sga=True, diff_sampling=diff, feature_dim=C, latent_dim=C)
        cfg = {k: v for k, v in cfg.items() if k not in ("ldecode_enabled", "ldecode_type", "norm_every", "decay_period", "temperature")}
        dec = LatentDecoder(**cfg)
        dec.temperature = tau
        with torch.no_grad():
            dec.layers[0].scale.copy_(torch.eye(C))
        torch.manual_seed(500 + si)
        w = (torch.rand(T, C) - 0.5) * 16
        w[:8, 0] = torch.tensor([0.0, 1.0, -3.0, 2.5, -2.5, 0.9999995, -0.0000005, 7.000001])   # integers, halves, clamp edges
        gout = torch.randn(T, C)
        w = w.clone().requires_grad_(True)
        torch.manual_seed(900 + si)
        u = torch.rand(T, C, 2)
        torch.manual_seed(900 + si)
        w_out = dec(w)
        (w_out * gout).sum().backward()
        p = "dec/" + name + "/"
        out[p + "meta"] = np.array([T, C, 1 if diff else 0], dtype=np.int32)
        out[p + "tau"] = np.array([tau], dtype=np.float64)
        out[p + "w"] = w.detach().numpy().copy()
        out[p + "u"] = u.numpy().copy()
        out[p + "w_hat"] = w_out.detach().numpy().copy()
        out[p + "gout"] = gout.numpy().copy()
        out[p + "grad_w"] = w.grad.numpy().copy()
    out["dec_cases"] = np.array([s[0] for s in specs])
    # ---- grid level: LatentGrid.interpolate with use_sga -------------------------------------------------------------
    torch.manual_seed(4242)
    dim, L, bw, rmin, rmax, C, F, tau = 2, 8, 10, 16, 128, 1, 1, 0.6
    dcfg = dec_cfg("single", C, True, "sq")
    dcfg.update(use_sga=True, diff_sampling=True)
    grid = lg.LatentGrid.from_geometric(feature_dim=F, num_lods=L, latent_dim=C, multiscale_type="cat", resolution_dim=dim,
                                        feature_std=0.1, codebook_bitwidth=bw, min_grid_res=rmin, max_grid_res=rmax,
                                        init_grid="uniform", conf_latent_decoder=dcfg, conf_entropy_reg=ent_cfg(2))
    grid.latent_dec.temperature = tau
    with torch.no_grad():
        grid.codebook.mul_(60.0)
        grid.latent_dec.div.data = torch.rand(C) * 2 + 0.5
        grid.latent_dec.layers[0].shift.normal_(0, 0.05)
    N = 1200
    coords = torch.rand(N, dim) * 2 - 1
    gout = torch.randn(N, L * F)
    torch.manual_seed(77)
    u = torch.rand(grid.codebook.shape[0], C, 2)
    torch.manual_seed(77)
    feats = grid.interpolate(coords, 0)
    grid.zero_grad()
    feats.backward(gout)
    p = "grid/"
    out[p + "meta"] = np.array([dim, L, bw, C, F], dtype=np.int32)
    out[p + "tau"] = np.array([tau], dtype=np.float64)
    out[p + "resolutions"] = np.array(grid.resolutions, dtype=np.int32)
    out[p + "codebook"] = grid.codebook.detach().numpy().copy()
    out[p + "u"] = u.numpy().copy()
    out[p + "coords"] = coords.numpy()
    out[p + "grad_out"] = gout.numpy()
    out[p + "feats"] = feats.detach().numpy()
    out[p + "grad_codebook"] = grid.codebook.grad.numpy().copy()
    d = grid.latent_dec
    out[p + "div"] = d.div.detach().numpy().copy()
    out[p + "scale"] = d.layers[0].scale.detach().numpy().copy()
    out[p + "shift"] = d.layers[0].shift.detach().numpy().copy()
    out[p + "grad_scale"] = d.layers[0].scale.grad.numpy().copy()
    out[p + "grad_shift"] = d.layers[0].shift.grad.numpy().copy()
    path = os.path.join(ROOT, "tests", "golden", "sga_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
