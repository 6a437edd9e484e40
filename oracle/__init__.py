"""CPU oracle of the latent hash-grid hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs
may import this package; the product (shacira_b200/) never does. See hashgrid_oracle.c and
latent_oracle.py for the reference file:line each function restates.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "hashgrid_oracle.c")
_LIB = os.path.join(_HERE, "liboracle_hashgrid.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        # -ffp-contract=off: FMAs only where the restatement asks for them (fmaf)
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-fPIC", "-ffp-contract=off", "-shared", "-o", _LIB, _SRC,
                               "-lm"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        L.oracle_hashgrid_corners.restype = None
        L.oracle_hashgrid_corners.argtypes = [ctypes.c_int, vp, i64, vp, i32, i32, vp, vp]
        L.oracle_hashgrid_forward.restype = i64
        L.oracle_hashgrid_forward.argtypes = [ctypes.c_int, vp, i64, vp, i64, vp, vp, i32, i32, i32, vp]
        L.oracle_hashgrid_backward.restype = i64
        L.oracle_hashgrid_backward.argtypes = [ctypes.c_int, vp, i64, vp, i64, vp, vp, i32, i32, i32, vp]
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.restype = None
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads():
    return int(lib().oracle_num_threads())


def use_all_cores():
    """All host cores for the timed CPU arm (torchrun sets OMP_NUM_THREADS=1 for its ranks). Returns the count."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_num_threads(n)
    try:
        import torch
        torch.set_num_threads(n)
    except Exception:
        pass
    return num_threads()


def level_layout(resolutions, bitwidth, dim):
    """(sizes, first_idx, total) as the grid constructors lay the table out (latent_grid.py:100-112)."""
    sizes = [min(2 ** bitwidth, int(r) ** dim) for r in resolutions]
    first = [0]
    for s in sizes[:-1]:
        first.append(first[-1] + s)
    return sizes, first, sum(sizes)


def geometric_resolutions(min_res, max_res, num_lods):
    """latent_grid.py:280-281."""
    b = np.exp((np.log(max_res) - np.log(min_res)) / (num_lods - 1))
    return [int(1 + np.floor(min_res * (b ** l))) for l in range(num_lods)]


def corners(coords, resolutions, bitwidth):
    coords = _f32(coords)
    n, dim = coords.shape
    res = _i32(resolutions)
    idx = np.empty((n, len(res), 1 << dim), dtype=np.int32)
    w = np.empty((n, len(res), 1 << dim), dtype=np.float32)
    lib().oracle_hashgrid_corners(dim, coords.ctypes.data, n, res.ctypes.data, len(res), bitwidth, idx.ctypes.data,
                                  w.ctypes.data)
    return idx, w


def forward(coords, codebook, first_idx, resolutions, bitwidth):
    coords, codebook = _f32(coords), _f32(codebook)
    n, dim = coords.shape
    T, F = codebook.shape
    res, first = _i32(resolutions), _i32(first_idx)
    feats = np.empty((n, len(res) * F), dtype=np.float32)
    lib().oracle_hashgrid_forward(dim, coords.ctypes.data, n, codebook.ctypes.data, T, first.ctypes.data,
                                  res.ctypes.data, len(res), bitwidth, F, feats.ctypes.data)
    return feats


def backward(coords, grad_output, table_rows, first_idx, resolutions, bitwidth, feature_dim):
    coords, grad_output = _f32(coords), _f32(grad_output)
    n, dim = coords.shape
    res, first = _i32(resolutions), _i32(first_idx)
    grad = np.empty((table_rows, feature_dim), dtype=np.float32)
    lib().oracle_hashgrid_backward(dim, coords.ctypes.data, n, grad_output.ctypes.data, table_rows, first.ctypes.data,
                                   res.ctypes.data, len(res), bitwidth, feature_dim, grad.ctypes.data)
    return grad
