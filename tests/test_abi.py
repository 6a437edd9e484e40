"""The C-ABI library loads and exports every symbol include/shacira_b200.h declares.
CPU only: argument validation runs before any CUDA call, so error paths are testable here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "shacira_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(shacira_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared_functions()
    for must in ("shacira_hashgrid_forward", "shacira_hashgrid_backward", "shacira_latent_forward",
                 "shacira_latent_backward", "shacira_entropy_bits", "shacira_symbol_histogram",
                 "shacira_latent_step_host", "shacira_ac_encode", "shacira_ac_decode"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    L = ctypes.CDLL(lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(L, name), name


def test_python_binding_covers_every_declared_symbol(lib):
    assert sorted(lib.SIGNATURES) == _declared_functions()


def test_abi_version(lib):
    assert lib.load().shacira_abi_version() == 1


def _levels(vals):
    return (ctypes.c_int32 * len(vals))(*vals)


def test_invalid_arguments_are_reported_not_crashed(lib):
    L = lib.load()
    res, first = _levels([17, 33]), _levels([0, 289])
    # bad dim
    rc = L.shacira_hashgrid_forward(4, None, 0, None, first, res, 2, 10, 2, None, None)
    assert rc == lib.ERR_INVALID_ARGUMENT and b"dim" in L.shacira_last_error()
    # too many levels
    rc = L.shacira_hashgrid_forward(2, None, 0, None, first, res, 99, 10, 2, None, None)
    assert rc == lib.ERR_INVALID_ARGUMENT
    # empty input is a successful no-op (the reference launches a zero-sized grid and fails; we do not)
    rc = L.shacira_hashgrid_forward(2, None, 0, None, first, res, 2, 10, 2, None, None)
    assert rc == lib.OK


def test_reference_int32_overflow_window_is_fenced(lib):
    """res >= 1291 with res^2 < T makes the reference's int32 res^3 wrap and take the
    dense branch out of bounds. The library refuses those levels instead of guessing."""
    L = lib.load()
    res, first = _levels([1483]), _levels([0])
    rc = L.shacira_hashgrid_forward(3, None, 0, None, first, res, 1, 22, 2, None, None)
    assert rc == lib.ERR_Q2_WINDOW
    res = _levels([1290])
    assert L.shacira_hashgrid_forward(3, None, 0, None, first, res, 1, 22, 2, None, None) == lib.OK
    # 2D never overflows at these sizes
    res = _levels([2049])
    assert L.shacira_hashgrid_forward(2, None, 0, None, first, res, 1, 22, 2, None, None) == lib.OK


def test_missing_library_fails_loudly(lib, monkeypatch):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libshacira_b200.so")
    with pytest.raises(RuntimeError, match="no CPU"):
        lib.load()


def test_cpu_tensors_are_rejected(lib):
    import torch
    with pytest.raises(lib.ShaciraError, match="CUDA tensor"):
        lib.hashgrid_forward(torch.zeros(4, 2), torch.zeros(289, 2), [0], [17], 10)
