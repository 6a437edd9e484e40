"""NeRF-shape fit (BASELINE cfg4) end to end with the synthetic sampler of SURVEY 8d: PSNR / size parity of this package's
3D latent grid against the same fit driven by the reference's OWN CUDA kernels (oracle/_ref).

Scene: an analytic radiance field in [-1, 1]^3 (a few soft coloured blobs). Every step draws RAYS rays through the cube,
takes K samples per ray (uniform along the chord inside the cube, jittered), queries

    feats = LatentGrid.interpolate(samples)                   3D, 16 levels 16 -> 2048, 2^19 rows, C = 1 -> F = 4
    density = softplus(MLP_d(feats)),  rgb = sigmoid(MLP_c(feats))        (plain PyTorch, wisp/models/nefs/nerf.py:218-233)
    pixel = exponential integration of (rgb, density * delta) per ray     (packed_rf_tracer.py:136-153; kernel or torch)

and minimises the L1 colour error plus lambda * bits / rows with the bit-rate loss evaluated the way the reference's
NeRF trainer does (is_val = pipeline.training, i.e. on round(w): multiview_trainer.py:110). SGA is off in
both arms (RNG-bound); RMSprop lr 1e-3 / grid lr 1e-2 (the effective nerf_lego.yaml values).

  --impl ours   shacira_b200.grids.LatentGrid (sorted lane-pair kernels, plan re-binned per step)
  --impl ref    LatentGrid.interpolate restated around the reference's kernels: table-side round + decode with torch,
                one launch per level (oracle/_ref)

Prints one JSON line: PSNR on a fixed set of test rays, latent bits (empirical entropy of round(w)), ms/step."""
import argparse
import copy
import json
import math
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

RAYS, K = 4096, 128
L, BW, C, F = 16, 19, 1, 4
DEC = dict(ldecode_enabled=True, ldecode_type="single", use_sga=False, diff_sampling=True, use_shift=True,
           ldecode_matrix="sq", latent_dim=C, norm="none", norm_every=10, ldec_std=1.0, decay_period=0.9, temperature=1.0)
ENT = dict(num_prob_layers=1, entropy_reg=1e-4, entropy_reg_end=1e-4, entropy_reg_sched="fix", noise_freq=1)
LATENT_SCALE = 20.0


def scene(x):
    """Analytic field: density [N] >= 0 and colour [N, 3] in [0, 1] at points x [N, 3]."""
    centres = torch.tensor([[0.3, 0.2, -0.1], [-0.4, -0.3, 0.25], [0.0, 0.5, 0.4], [0.45, -0.45, -0.4]], device=x.device)
    cols = torch.tensor([[0.9, 0.2, 0.1], [0.1, 0.8, 0.3], [0.2, 0.3, 0.9], [0.9, 0.8, 0.2]], device=x.device)
    rad = torch.tensor([0.35, 0.3, 0.25, 0.3], device=x.device)
    d2 = ((x[:, None, :] - centres[None]) ** 2).sum(-1)
    w = torch.exp(-d2 / (2 * (rad * 0.5) ** 2))
    density = 12.0 * w.sum(1)
    colour = (w[..., None] * cols[None]).sum(1) / (w.sum(1, keepdim=True) + 1e-6)
    ripple = 0.1 * torch.sin(9.0 * x).prod(-1, keepdim=True)
    return density, (colour + ripple).clamp(0, 1)


def make_rays(n, gen, dev):
    """Rays through the cube: origin on a sphere of radius 2.5, aimed at a random point of the cube."""
    o = torch.randn(n, 3, generator=gen)
    o = 2.5 * o / o.norm(dim=1, keepdim=True)
    tgt = (torch.rand(n, 3, generator=gen) - 0.5) * 1.2
    d = tgt - o
    d = d / d.norm(dim=1, keepdim=True)
    o, d = o.to(dev), d.to(dev)
    inv = 1.0 / d
    t0, t1 = (-1 - o) * inv, (1 - o) * inv
    tn = torch.minimum(t0, t1).amax(1)
    tf = torch.maximum(t0, t1).amin(1)
    return o, d, tn, tf


def sample(o, d, tn, tf, gen):
    n = o.shape[0]
    u = (torch.arange(K, device=o.device)[None] + torch.rand(n, K, generator=gen).to(o.device)) / K
    t = tn[:, None] + (tf - tn)[:, None] * u
    pts = (o[:, None, :] + t[..., None] * d[:, None, :]).reshape(-1, 3).clamp(-1, 1)
    delta = ((tf - tn) / K)[:, None].expand(n, K).reshape(-1)
    return pts.contiguous(), delta.contiguous()


def integrate(rgb, tau, n):
    """exclusive-cumsum transmittance (packed_rf_tracer.py:136-153 restated; every ray has exactly K samples here)."""
    tau = tau.reshape(n, K)
    T = torch.exp(-(torch.cumsum(tau, 1) - tau))
    w = T * (1 - torch.exp(-tau))
    return (w[..., None] * rgb.reshape(n, K, 3)).sum(1)


def target_pixels(o, d, tn, tf, gen):
    pts, delta = sample(o, d, tn, tf, gen)
    dens, col = scene(pts)
    return integrate(col, dens * delta, o.shape[0])


class RefGrid3D(nn.Module):
    """The reference's LatentGrid.interpolate (latent_grid.py:355-370) around the reference's own 3D kernels."""

    def __init__(self, ours):
        super().__init__()
        from oracle import build_ref
        self.ref = build_ref.load()
        assert self.ref is not None, "oracle/_ref/wisp_ref_ops.so missing"
        self.codebook = nn.Parameter(ours.codebook.detach().clone())
        self.latent_dec = copy.deepcopy(ours.latent_dec)
        self.prob_model = copy.deepcopy(ours.prob_model)
        self.register_buffer("first_idx", ours.codebook_lod_first_idx.clone())
        self.resolutions, self.bw = list(ours.resolutions), ours.codebook_bitwidth
        ref, me, res, bw = self.ref, self, self.resolutions, self.bw

        class Fn(torch.autograd.Function):  # wisp/ops/grid.py:69-111
            @staticmethod
            def forward(ctx, coords, table):
                ctx.save_for_backward(coords, table)
                return ref.hashgrid_interpolate_cuda(coords, table, me.first_idx, res, bw)

            @staticmethod
            def backward(ctx, g):
                coords, table = ctx.saved_tensors
                return None, ref.hashgrid_interpolate_backward_cuda(coords, g.contiguous(), table, me.first_idx, res, bw,
                                                                    table.shape[1], False)
        self.fn = Fn

    def interpolate(self, coords, lod_idx):
        return self.fn.apply(coords, self.latent_dec(self.codebook))

    def ent_loss(self, idx, is_val=False):
        w = torch.round(self.codebook) if is_val else self.codebook
        prob = self.prob_model(w + 0.5) - self.prob_model(w - 0.5)
        bits = torch.sum(torch.clamp(-1.0 * torch.log(prob + 1e-10) / np.log(2.0), 0, 50))
        return bits / self.codebook.shape[0], bits


def psnr(a, b):
    return -10.0 * math.log10(max(float(((a - b) ** 2).mean()), 1e-12))


def fit(seed, impl, steps, dev):
    from shacira_b200.grids import LatentGrid
    torch.manual_seed(seed)
    grid = LatentGrid.from_geometric(feature_dim=F, num_lods=L, latent_dim=C, multiscale_type="cat", resolution_dim=3,
                                     feature_std=0.1, codebook_bitwidth=BW, min_grid_res=16, max_grid_res=2048,
                                     init_grid="uniform", conf_latent_decoder=dict(DEC), conf_entropy_reg=dict(ENT))
    head_d = nn.Sequential(nn.Linear(L * F, 64), nn.ReLU(), nn.Linear(64, 1))
    head_c = nn.Sequential(nn.Linear(L * F, 64), nn.ReLU(), nn.Linear(64, 3))
    with torch.no_grad():
        grid.codebook.mul_(LATENT_SCALE)       # rounding is non-trivial from step 0 (SGA is off in both arms)
        grid.latent_dec.layers[0].scale.mul_(0.1)
    grid, head_d, head_c = grid.to(dev), head_d.to(dev), head_c.to(dev)
    if impl == "ref":
        grid = RefGrid3D(grid).to(dev)
    params = [dict(params=list(head_d.parameters()) + list(head_c.parameters()), lr=1e-3),
              dict(params=[grid.codebook], lr=1e-2),
              dict(params=[p for p in grid.latent_dec.parameters() if p.requires_grad], lr=1e-3),
              dict(params=list(grid.prob_model.parameters()), lr=1e-4)]
    opt = torch.optim.RMSprop(params)
    gen = torch.Generator().manual_seed(1000 + seed)
    test_gen = torch.Generator().manual_seed(77)
    to, td, ttn, ttf = make_rays(RAYS, test_gen, dev)
    test_px = target_pixels(to, td, ttn, ttf, test_gen)

    def render(o, d, tn, tf, g):
        pts, delta = sample(o, d, tn, tf, g)
        feats = grid.interpolate(pts, 0)
        dens = torch.nn.functional.softplus(head_d(feats)).squeeze(1)
        col = torch.sigmoid(head_c(feats))
        return integrate(col, dens * delta, o.shape[0])

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(steps):
        o, d, tn, tf = make_rays(RAYS, gen, dev)
        gt = target_pixels(o, d, tn, tf, gen)
        opt.zero_grad(set_to_none=True)
        pred = render(o, d, tn, tf, gen)
        avg_bits, _ = grid.ent_loss(it, is_val=True)         # the NeRF trainer's (inverted) flag: SURVEY Q8
        loss = torch.abs(pred - gt).mean() + 1e-4 * avg_bits
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    with torch.no_grad():
        eval_gen = torch.Generator().manual_seed(78)
        pred = render(to, td, ttn, ttf, eval_gen)
        q = torch.round(grid.codebook.detach()[:, 0]).long()
        _, counts = torch.unique(q, return_counts=True)
        p = counts / counts.sum()
        bits = float(torch.sum(torch.clamp(-torch.log(p + 1e-10) / np.log(2.0), 0, 1000) * counts))
    return dict(seed=seed, impl=impl, psnr=psnr(pred, test_px), latent_bits=bits, mbytes=bits / 8e6, ms_per_step=dt / steps * 1e3,
                loss=float(loss.detach()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours", choices=["ours", "ref"])
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    print(json.dumps(fit(args.seed, args.impl, args.steps, torch.device("cuda", 0))), flush=True)


if __name__ == "__main__":
    main()
