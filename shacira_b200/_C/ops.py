"""The four entry points the reference registers as `wisp._C.ops`
(wisp/csrc/bindings.cpp:23-27; declarations wisp/csrc/ops/hashgrid_interpolate.h:18-50),
same names, argument order and return values, implemented on the B200 C-ABI library.

Differences from the reference, all deliberate:
  * fp32 only (the reference also instantiates double/half); other dtypes raise.
  * inputs are validated (device, dtype, shape); the reference checks nothing.
  * `require_grad_coords` is accepted and ignored: the reference computes grad_coords with
    wrong indices and never returns it (hashgrid_interpolate.cpp:182).
  * all levels run in ONE kernel launch instead of one launch per level.
"""
import torch

from .. import _lib

_host_cache = {}


def _host_ints(t):
    """Host copy of a small int tensor (first_idx), cached on (ptr, version) to avoid a sync per call."""
    if not isinstance(t, torch.Tensor):
        return tuple(int(v) for v in t)
    key = (t.data_ptr(), t._version, t.numel(), str(t.device))
    got = _host_cache.get(key)
    if got is None:
        got = tuple(int(v) for v in t.detach().cpu().tolist())
        if len(_host_cache) > 256:
            _host_cache.clear()
        _host_cache[key] = got
    return got


def hashgrid_interpolate_cuda(coords, codebook, codebook_first_idx, resolution, codebook_bitwidth):
    return _lib.hashgrid_forward(coords, codebook, _host_ints(codebook_first_idx), list(resolution),
                                 int(codebook_bitwidth))


def hashgrid_interpolate_backward_cuda(coords, grad_output, codebook, codebook_first_idx, resolution,
                                       codebook_bitwidth, feature_dim, require_grad_coords):
    return _lib.hashgrid_backward(coords, grad_output, _host_ints(codebook_first_idx), list(resolution),
                                  int(codebook_bitwidth), int(feature_dim), codebook.shape[0])


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
