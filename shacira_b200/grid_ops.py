"""Host-side mirror of the reference's `wisp/ops/grid.py` (autograd Functions + wrappers).

`hashgrid` / `hashgrid2d` keep the reference signatures (wisp/ops/grid.py:113-131,178-196),
including the arguments the reference accepts and ignores (`lod_idx`, `codebook_sizes`,
SURVEY Q9) and the exception for odd feature dims (:75,:140, Q10). `latent_hashgrid` is the
fused quantize -> gather -> lerp -> decode op that `LatentGrid.interpolate` uses instead of
`latent_dec(codebook)` followed by `hashgrid*()`.
"""
import torch

from . import _lib
from ._C import ops as _ops
from ._C.ops import _host_ints

import os
from collections import OrderedDict

# ---- plan cache --------------------------------------------------------------------------------
# The tiled kernels need the points binned by spatial tile. The reference's API hands the op a bare
# coords tensor every step, so plans are cached on the tensor's identity: (storage pointer, version
# counter, shape). The cache keeps a reference to the tensor, so its storage cannot be freed and
# handed to different data while the entry lives, and any in-place write bumps the version counter.
# An image fit calls interpolate() with the same coords tensor for every step (static coordinates,
# image_trainer.py:234-266): one plan for the whole fit. A workload with fresh coordinates every
# step (NeRF samples) misses the cache and pays the ~3-kernel plan build each step.
# Crossovers measured on the B200 (benchmarks/crossover.py, profiles/r02h_crossover.jsonl). 2D, cfg2 grid: fwd + bwd tiled
# 67 us against 90 us point-parallel at 2^16 points (83 us when the plan is re-binned every step), but 115 / 147 us
# against 94 us at 2^14 / 2^15 -- a 64-tile plan leaves most SMs idle. 3D, cfg4 grid: sorted 376 us (415 us with the
# per-step re-binning) against 436 us unsorted at 2^19 samples; below that the re-binning costs more than the sort saves.
PLAN_MIN_POINTS = int(os.environ.get("SHACIRA_PLAN_MIN_POINTS", "65536"))
PLAN_CACHE_SIZE = int(os.environ.get("SHACIRA_PLAN_CACHE", "8"))
_plans = OrderedDict()
plan_stats = {"hits": 0, "builds": 0}


PLAN_MIN_POINTS_3D = int(os.environ.get("SHACIRA_PLAN_MIN_POINTS_3D", "393216"))


class _PlanLease:
    """Held by an autograd ctx for as long as its graph lives: the plan cache does not re-bin a leased plan for other
    coordinates (a second backward under retain_graph, or a backward that never runs, keep the lease -- it is released
    when the ctx is collected, not when backward() happens to be called)."""

    def __init__(self, plan):
        self.plan = plan
        plan.users += 1

    def __del__(self):
        try:
            self.plan.users = max(0, self.plan.users - 1)
        except Exception:
            pass


def plan_for(coords):
    """Tile plan for `coords` ([N, 2|3] float32 contiguous CUDA) or None when planning is disabled / not worth it.
    3D sample sets (new coordinates every step) re-bin a recycled plan: ~40 us at the NeRF batch, paid back by the
    sorted kernels from ~2^16 samples on (profiles/r02b_probe3d.jsonl)."""
    if os.environ.get("SHACIRA_DISABLE_PLAN") or coords.shape[0] < PLAN_MIN_POINTS:
        return None
    if coords.shape[1] == 3 and (os.environ.get("SHACIRA_DISABLE_PLAN_3D") or coords.shape[0] < PLAN_MIN_POINTS_3D):
        return None
    key = (coords.data_ptr(), coords._version, tuple(coords.shape), coords.device.index)
    plan = _plans.get(key)
    if plan is not None:
        _plans.move_to_end(key)
        plan_stats["hits"] += 1
        return plan
    plan = None
    if len(_plans) >= PLAN_CACHE_SIZE:
        # recycle the least recently used plan: its device allocation is reused (no cudaMalloc on the hot path
        # of workloads whose coordinates change every step)
        _, old = _plans.popitem(last=False)
        if old.users > 0:
            pass        # a pending backward still needs it: leave it to its owner (freed when the graph dies)
        elif old.device == coords.device:
            plan = old.rebuild(coords)
        else:
            old.close()
    if plan is None:
        plan = _lib.Plan(coords)
    plan_stats["builds"] += 1
    _plans[key] = plan
    return plan


def clear_plans():
    while _plans:
        _, old = _plans.popitem()
        old.close()


_amp_fwd = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_amp_bwd = torch.amp.custom_bwd(device_type="cuda")


class _ShapeOnly:
    """Stands in for the codebook in the backward entry point, which only reads its shape (as the reference's does)."""

    def __init__(self, rows, feature_dim):
        self.shape = (rows, feature_dim)


class HashGridInterpolate(torch.autograd.Function):
    """wisp/ops/grid.py:69-111. fp32 compute also under autocast (the reference casts to half)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, coords, resolutions, codebook_bitwidth, lod_idx, codebook, codebook_sizes, codebook_first_idx):
        if codebook[0].shape[-1] % 2 == 1:
            raise Exception("The codebook feature dimension needs to be a multiple of 2.")
        feats_out = _ops.hashgrid_interpolate_cuda(coords.float().contiguous(), codebook, codebook_first_idx,
                                                   resolutions, codebook_bitwidth)
        ctx.save_for_backward(coords, codebook_first_idx)
        ctx.resolutions = resolutions
        ctx.codebook_bitwidth = codebook_bitwidth
        ctx.table_shape = tuple(codebook.shape)
        return feats_out

    @staticmethod
    @_amp_bwd
    def backward(ctx, grad_output):
        coords, first_idx = ctx.saved_tensors
        rows, feature_dim = ctx.table_shape
        # through the drop-in entry point: large 2D batches reuse the forward's cached tile plan
        grad_codebook = _ops.hashgrid_interpolate_backward_cuda(coords.float().contiguous(), grad_output.contiguous(),
                                                                _ShapeOnly(rows, feature_dim), first_idx,
                                                                ctx.resolutions, ctx.codebook_bitwidth, feature_dim,
                                                                False)
        return (None, None, None, None, grad_codebook, None, None)


class HashGridInterpolate2D(HashGridInterpolate):
    """wisp/ops/grid.py:135-176 (one kernel family serves both dimensions here)."""


def hashgrid(coords, resolutions, codebook_bitwidth, lod_idx, codebook, codebook_sizes, codebook_first_idx):
    """3D hash-grid query + trilinear interpolation. coords [batch, 3] -> [batch, F * num_lods]."""
    batch, dim = coords.shape
    feats = HashGridInterpolate.apply(coords.contiguous(), resolutions, codebook_bitwidth, lod_idx, codebook,
                                      codebook_sizes, codebook_first_idx)
    return feats.reshape(batch, codebook.shape[1] * len(resolutions))


def hashgrid2d(coords, resolutions, codebook_bitwidth, lod_idx, codebook, codebook_sizes, codebook_first_idx):
    """2D hash-grid query + bilinear interpolation. coords [batch, 2] -> [batch, F * num_lods]."""
    batch, dim = coords.shape
    feats = HashGridInterpolate2D.apply(coords.contiguous(), resolutions, codebook_bitwidth, lod_idx, codebook,
                                        codebook_sizes, codebook_first_idx)
    return feats.reshape(batch, codebook.shape[1] * len(resolutions))


class _HashGridAnyF(torch.autograd.Function):
    """Plain interpolate without the even-F restriction (the kernels handle F = 1 natively)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, coords, codebook, first_idx, resolutions, bitwidth):
        ctx.save_for_backward(coords)
        ctx.meta = (first_idx, tuple(resolutions), bitwidth, tuple(codebook.shape))
        return _lib.hashgrid_forward(coords, codebook, first_idx, resolutions, bitwidth)

    @staticmethod
    @_amp_bwd
    def backward(ctx, grad_output):
        (coords,) = ctx.saved_tensors
        first_idx, resolutions, bitwidth, (rows, F) = ctx.meta
        g = _lib.hashgrid_backward(coords, grad_output.contiguous(), first_idx, resolutions, bitwidth, F, rows)
        return (None, g, None, None, None)


def hashgrid_any(coords, codebook, first_idx, resolutions, bitwidth):
    return _HashGridAnyF.apply(coords.float().contiguous(), codebook, _host_ints(first_idx), list(resolutions),
                               int(bitwidth))


class LatentHashGrid(torch.autograd.Function):
    """Fused LatentGrid.interpolate for affine latent decoders (latent_grid.py:359-368 +
    basic_latent_decoder.py:85-95,192-194): feats = lerp(round?(latents)) @ A + shift."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, coords, latents, A, shift, first_idx, resolutions, bitwidth, round_flag):
        F = A.shape[2]
        need_dec = bool(ctx.needs_input_grad[2] or ctx.needs_input_grad[3])
        dim3 = coords.shape[1] == 3
        if dim3:  # sorted lane-pair kernels: any level count, latent_dim 1 or 2
            plan = plan_for(coords) if latents.shape[1] in (1, 2) and latents.data_ptr() % 16 == 0 else None
        else:
            plan = plan_for(coords) if len(resolutions) % 4 == 0 else None  # tiled kernels unroll 4 levels
        if plan is not None and dim3:
            feats, z = _lib.latent_forward_planned_z(plan, latents, first_idx, resolutions, bitwidth, A, shift, F,
                                                     round_flag, save_z=need_dec)
            ctx.save_for_backward(coords, A.detach(), z if z is not None else coords.new_empty(0))
        elif plan is not None:
            feats = _lib.latent_forward_planned(plan, latents, first_idx, resolutions, bitwidth, A, shift, F, round_flag)
            # the tiled backward recomputes the interpolation from the latents: nothing extra is written here
            ctx.save_for_backward(coords, A.detach(), latents.detach() if need_dec else coords.new_empty(0))
        else:
            feats, z = _lib.latent_forward(coords, latents, first_idx, resolutions, bitwidth, A, shift, F, round_flag,
                                           save_z=need_dec)
            ctx.save_for_backward(coords, A.detach(), z if z is not None else coords.new_empty(0))
        ctx.plan = plan
        ctx.lease = _PlanLease(plan) if (plan is not None and any(ctx.needs_input_grad)) else None
        ctx.round_flag = round_flag
        ctx.meta = (first_idx, tuple(resolutions), bitwidth, tuple(latents.shape), F, need_dec, shift is not None,
                    A.shape[0])
        return feats

    @staticmethod
    @_amp_bwd
    def backward(ctx, grad_output):
        coords, A, z = ctx.saved_tensors
        first_idx, resolutions, bitwidth, (rows, C), F, need_dec, has_shift, nA = ctx.meta
        if ctx.plan is not None and coords.shape[1] == 3:
            gl, gA, gS = _lib.latent_backward_planned_z(ctx.plan, grad_output.contiguous(), z if need_dec else None,
                                                        first_idx, resolutions, bitwidth, A, C, F, rows, need_dec)
        elif ctx.plan is not None:
            gl, gA, gS = _lib.latent_backward_planned(ctx.plan, grad_output.contiguous(), z if need_dec else None,
                                                      first_idx, resolutions, bitwidth, A, C, F, rows,
                                                      ctx.round_flag, need_dec)
        else:
            gl, gA, gS = _lib.latent_backward(coords, grad_output.contiguous(), z if need_dec else None, first_idx,
                                              resolutions, bitwidth, A, C, F, rows, need_dec)
        if need_dec and nA == 1:
            gA = gA.sum(0, keepdim=True)
            gS = gS.sum(0, keepdim=True)
        if not ctx.needs_input_grad[1]:
            gl = None
        return (None, gl, gA if need_dec else None, gS if (need_dec and has_shift) else None, None, None, None, None)


def latent_hashgrid(coords, latents, A, shift, first_idx, resolutions, bitwidth, round_flag=True):
    """coords [N, 2|3]; latents [T, C]; A [1|L, C, F]; shift [1|L, F] | None -> feats [N, L*F]."""
    return LatentHashGrid.apply(coords.float().contiguous(), latents, A, shift, _host_ints(first_idx),
                                list(resolutions), int(bitwidth), bool(round_flag))


class SGAQuantize(torch.autograd.Function):
    """Fused SGA sample of the latents (basic_latent_decoder.py:183-191): one table-side kernel produces w_hat and
    d w_hat / d w; backward is one multiply."""

    @staticmethod
    def forward(ctx, latents, temperature, diff_sampling, uniforms, seed):
        w_hat, dw = _lib.sga_quantize(latents, temperature, diff_sampling, uniforms=uniforms, seed=seed,
                                      want_dw=ctx.needs_input_grad[0])
        if dw is not None:
            ctx.save_for_backward(dw)
        return w_hat

    @staticmethod
    def backward(ctx, grad):
        (dw,) = ctx.saved_tensors
        return grad * dw, None, None, None, None


def sga_quantize(latents, temperature, diff_sampling, uniforms=None, seed=None):
    """latents [T, C] (CUDA) -> SGA sample w_hat [T, C]. `uniforms` [T, C, 2]: injected U(0,1) draws (parity runs);
    otherwise the noise is drawn inside the kernel from `seed` (default: taken from torch's CPU generator, so
    torch.manual_seed makes runs repeatable)."""
    if seed is None and uniforms is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return SGAQuantize.apply(latents, temperature, bool(diff_sampling), uniforms, seed or 0)


class EntropyBits(torch.autograd.Function):
    """Fused LatentGrid.ent_loss body (latent_grid.py:132-135): total bits of the factorized density.
    The kernel produces value and gradients in one pass; backward only scales them."""

    @staticmethod
    def forward(ctx, latents, noise, params, num_layers, first_idx):
        bits, gl, gp = _lib.entropy_bits(latents, noise, params, num_layers, first_idx, want_grads=True)
        ctx.save_for_backward(gl, gp)
        ctx.per_level = bits[1:].to(torch.float32)
        return bits[0].to(torch.float32)

    @staticmethod
    def backward(ctx, grad_total):
        gl, gp = ctx.saved_tensors
        return (gl * grad_total if ctx.needs_input_grad[0] else None, None,
                gp * grad_total if ctx.needs_input_grad[2] else None, None, None)


def entropy_bits(latents, noise, params, num_layers, first_idx=None):
    fi = _host_ints(first_idx) if first_idx is not None else None
    return EntropyBits.apply(latents, noise, params, int(num_layers), fi)


class FusedMLPMSE(torch.autograd.Function):
    """decoder MLP (Linear-ReLU-Linear-ReLU-Linear) + ((pred - target)**2).mean() in one
    kernel that also produces every gradient; backward only scales them by the upstream scalar."""

    @staticmethod
    def forward(ctx, features, target, W1, b1, W2, b2, W3, b3, want_pred):
        loss, gx, pred, packed = _lib.mlp_mse_step(features, target, W1, b1, W2, b2, W3, b3, want_pred)
        ctx.save_for_backward(gx, packed)
        ctx.dims = (features.shape[1], W1.shape[0], W3.shape[0])
        if pred is None:
            pred = features.new_empty(0)
        ctx.mark_non_differentiable(pred)
        return loss, pred

    @staticmethod
    def backward(ctx, grad_loss, _grad_pred):
        gx, packed = ctx.saved_tensors
        need = ctx.needs_input_grad
        # two kernels in total: the feature gradient and ALL weight gradients (packed) times the upstream scalar
        g = _lib.split_mlp_grads(packed * grad_loss, *ctx.dims) if any(need[2:8]) else [None] * 6
        return (gx * grad_loss if need[0] else None, None, g[0] if need[2] else None, g[1] if need[3] else None,
                g[2] if need[4] else None, g[3] if need[5] else None, g[4] if need[6] else None,
                g[5] if need[7] else None, None)


def mlp_mse_loss(features, target, mlp, want_pred=False):
    """`mlp`: nn.Sequential(Linear(in,16), ReLU, Linear(16,16), ReLU, Linear(16,3)) (the reference's image decoder).
    Returns (mse loss, pred [N,3] or an empty tensor)."""
    l1, l2, l3 = mlp[0], mlp[2], mlp[4]
    return FusedMLPMSE.apply(features.contiguous(), target.contiguous(), l1.weight, l1.bias, l2.weight, l2.bias,
                             l3.weight, l3.bias, bool(want_pred))
