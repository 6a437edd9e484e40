"""The four entry points the reference registers as `wisp._C.ops`
(wisp/csrc/bindings.cpp:23-27; declarations wisp/csrc/ops/hashgrid_interpolate.h:18-50),
same names, argument order and return values, implemented on the B200 C-ABI library.

Differences from the reference, all deliberate:
  * fp32 arithmetic. The reference instantiates double / float / half (AT_DISPATCH_FLOATING_TYPES_AND_HALF,
    hashgrid_interpolate2d_cuda.cu:115,251) and its autograd Functions cast the inputs to half under autocast
    (wisp/ops/grid.py:73,138 -- AMP is ON by default in app/nerf). Half / bfloat16 / double tables and gradients are
    accepted here: they are up-cast, the kernels compute in fp32 (the half kernels of the reference also ACCUMULATE in
    half), and the result comes back in the table's dtype like the reference's `at::empty(..., codebook.options())`.
  * inputs are validated (device, dtype, shape); the reference checks nothing.
  * `require_grad_coords` is accepted and ignored: the reference computes grad_coords with
    wrong indices and never returns it (hashgrid_interpolate.cpp:182).
  * all levels run in ONE kernel launch instead of one launch per level.
  * large 2D batches take the tiled fast path (spatial plan cached per coordinate tensor, shared-memory node staging,
    fixed-point shared-memory accumulation): a plain table is the latent grid with C = F, identity decoder and no
    rounding, so the same kernels serve it. Forward values are equal to the point-parallel kernel's (x*1 + y*0
    is exact); 3D and small batches stay on the point-parallel kernels.
"""
import torch

from .. import _lib

_host_cache = {}   # key -> (the tensor itself, its host copy)


def _host_ints(t):
    """Host copy of a small int tensor (first_idx: the reference passes it as a DEVICE tensor on every call), cached to
    avoid a device synchronisation per call. The entry HOLDS the tensor: its storage cannot be freed and handed to
    another tensor while the entry lives, so (pointer, version) identifies the contents -- a key without the reference
    goes stale as soon as the caching allocator recycles the address for a different first_idx."""
    if not isinstance(t, torch.Tensor):
        return tuple(int(v) for v in t)
    key = (t.data_ptr(), t._version, t.numel(), str(t.device))
    got = _host_cache.get(key)
    if got is None or got[0].data_ptr() != t.data_ptr():
        if len(_host_cache) >= 64:
            _host_cache.clear()
        got = (t.detach(), tuple(int(v) for v in t.detach().cpu().tolist()))
        _host_cache[key] = got
    return got[1]


_eye_cache = {}


def _identity_decoder(F, device):
    key = (F, str(device))
    A = _eye_cache.get(key)
    if A is None:
        A = torch.eye(F, dtype=torch.float32, device=device).unsqueeze(0).contiguous()
        _eye_cache[key] = A
    return A


def _tiled_plan(coords, feature_dim, resolution):
    """The cached tile plan of `coords` when the tiled kernels cover this call, else None."""
    if not (isinstance(coords, torch.Tensor) and coords.is_cuda and coords.dtype == torch.float32 and coords.dim() == 2
            and coords.shape[1] == 2 and coords.is_contiguous()):
        return None
    if feature_dim not in (1, 2, 4) or len(resolution) % 4:
        return None
    from .. import grid_ops
    return grid_ops.plan_for(coords)


_FLOATS = (torch.float32, torch.float16, torch.bfloat16, torch.float64)


def _f32(t, name):
    """fp32 view of a floating tensor (AMP hands these entry points half tables and gradients)."""
    if t.dtype not in _FLOATS:
        raise _lib.ShaciraError(_lib.ERR_UNSUPPORTED, "%s must be a floating tensor, got %s" % (name, t.dtype))
    return t if t.dtype == torch.float32 else t.float()


def hashgrid_interpolate_cuda(coords, codebook, codebook_first_idx, resolution, codebook_bitwidth):
    first, res, bw = _host_ints(codebook_first_idx), list(resolution), int(codebook_bitwidth)
    out_dtype = codebook.dtype
    coords, codebook = _f32(coords, "coords"), _f32(codebook, "codebook")
    plan = _tiled_plan(coords, codebook.shape[1] if codebook.dim() == 2 else 0, res)
    if plan is not None:
        F = codebook.shape[1]
        feats = _lib.latent_forward_planned(plan, codebook, first, res, bw, _identity_decoder(F, codebook.device), None,
                                            F, False)
    else:
        feats = _lib.hashgrid_forward(coords, codebook, first, res, bw)
    return feats if out_dtype == torch.float32 else feats.to(out_dtype)


def hashgrid_interpolate_backward_cuda(coords, grad_output, codebook, codebook_first_idx, resolution,
                                       codebook_bitwidth, feature_dim, require_grad_coords):
    first, res, bw = _host_ints(codebook_first_idx), list(resolution), int(codebook_bitwidth)
    F, rows = int(feature_dim), codebook.shape[0]
    out_dtype = getattr(codebook, "dtype", torch.float32)     # the reference returns zeros_like(codebook) + atomics
    coords, grad_output = _f32(coords, "coords"), _f32(grad_output, "grad_output")
    if tuple(grad_output.shape) != (coords.shape[0], len(res) * F):
        raise _lib.ShaciraError(_lib.ERR_INVALID_ARGUMENT, "grad_output must be [N, L*F] = %s, got %s" % (
            (coords.shape[0], len(res) * F), tuple(grad_output.shape)))
    plan = _tiled_plan(coords, F, res)
    if plan is not None:
        g = _lib.latent_backward_planned(plan, grad_output, None, first, res, bw,
                                         _identity_decoder(F, grad_output.device), F, F, rows, False, False)[0]
    else:
        g = _lib.hashgrid_backward(coords, grad_output, first, res, bw, F, rows)
    return g if out_dtype == torch.float32 else g.to(out_dtype)


def hashgrid_interpolate2d_cuda(coords, codebook, codebook_first_idx, resolution, codebook_bitwidth):
    if coords.shape[-1] != 2:
        raise _lib.ShaciraError(_lib.ERR_INVALID_ARGUMENT, "hashgrid_interpolate2d_cuda needs coords [N, 2]")
    return hashgrid_interpolate_cuda(coords, codebook, codebook_first_idx, resolution, codebook_bitwidth)


def hashgrid_interpolate2d_backward_cuda(coords, grad_output, codebook, codebook_first_idx, resolution,
                                         codebook_bitwidth, feature_dim, require_grad_coords):
    if coords.shape[-1] != 2:
        raise _lib.ShaciraError(_lib.ERR_INVALID_ARGUMENT, "hashgrid_interpolate2d_backward_cuda needs coords [N, 2]")
    return hashgrid_interpolate_backward_cuda(coords, grad_output, codebook, codebook_first_idx, resolution,
                                              codebook_bitwidth, feature_dim, require_grad_coords)
