"""The reference's OWN CUDA kernels (oracle/_ref, compiled for sm_100a from the files under
/root/reference by oracle/build_ref.py) against the C oracle and against this package.
This is what pins the oracle: the reference ships no CPU path, tests or golden vectors."""
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import make_case, rel_err

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    from oracle import build_ref
    build_ref.build()            # no-op on the GPU box (no /root/reference there): uses the shipped .so
    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref/wisp_ref_ops.so not present")
    return mod


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


REF_CASES = [(2, 16, 14, 16, 512, 5000, 2, "uniform"), (2, 16, 16, 16, 512, 5000, 2, "pixels"),
             (2, 16, 16, 16, 512, 3000, 4, "arbitrary"), (3, 16, 19, 16, 2048, 5000, 2, "arbitrary"),
             (3, 16, 19, 16, 2048, 2000, 4, "uniform"), (3, 24, 19, 16, 512, 2000, 2, "arbitrary")]


@pytest.mark.parametrize("dim,L,bw,rmin,rmax,n,F,kind", REF_CASES)
def test_reference_kernels_vs_oracle_and_ours(lib, ref, dim, L, bw, rmin, rmax, n, F, kind):
    c = make_case(dim, L, bw, rmin, rmax, n, F, seed=31 + dim + F, coord_kind=kind)
    coords, table, gout = _dev(c["coords"]), _dev(c["table"]), _dev(c["grad_out"])
    first_dev = torch.tensor(c["first_idx"], dtype=torch.int32, device="cuda")
    fwd = ref.hashgrid_interpolate2d_cuda if dim == 2 else ref.hashgrid_interpolate_cuda
    bwd = ref.hashgrid_interpolate2d_backward_cuda if dim == 2 else ref.hashgrid_interpolate_backward_cuda
    f_ref = fwd(coords, table, first_dev, c["resolutions"], bw).cpu().numpy()
    f_orc = oracle.forward(c["coords"], c["table"], c["first_idx"], c["resolutions"], bw)
    f_our = lib.hashgrid_forward(coords, table, c["first_idx"], c["resolutions"], bw).cpu().numpy()
    # the oracle restates the reference kernel's arithmetic, FMA order included: identical bits expected
    exact = np.array_equal(f_ref.view(np.uint32), f_orc.view(np.uint32))
    print("ref-vs-oracle forward bit-exact:", exact, "max rel", rel_err(f_orc, f_ref))
    assert rel_err(f_orc, f_ref) <= 1e-6
    assert rel_err(f_our, f_ref) <= 1e-5
    g_ref = bwd(coords, gout, table, first_dev, c["resolutions"], bw, F, False).cpu().numpy()
    g_orc = oracle.backward(c["coords"], c["grad_out"], c["T"], c["first_idx"], c["resolutions"], bw, F)
    g_our = lib.hashgrid_backward(coords, gout, c["first_idx"], c["resolutions"], bw, F, c["T"]).cpu().numpy()
    assert rel_err(g_orc, g_ref) <= 1e-4
    assert rel_err(g_our, g_ref) <= 1e-4


def test_reference_forward_bit_exact_with_oracle(lib, ref):
    """Strong pin: the C oracle reproduces the reference kernel's float results bit for bit."""
    bad = 0
    for dim, bw, rmax in ((2, 16, 512), (3, 19, 2048)):
        c = make_case(dim, 16, bw, 16, rmax, 20000, 2, seed=77, coord_kind="arbitrary")
        first_dev = torch.tensor(c["first_idx"], dtype=torch.int32, device="cuda")
        fwd = ref.hashgrid_interpolate2d_cuda if dim == 2 else ref.hashgrid_interpolate_cuda
        f_ref = fwd(_dev(c["coords"]), _dev(c["table"]), first_dev, c["resolutions"], bw).cpu().numpy()
        f_orc = oracle.forward(c["coords"], c["table"], c["first_idx"], c["resolutions"], bw)
        bad += int((f_ref.view(np.uint32) != f_orc.view(np.uint32)).sum())
    assert bad == 0, "%d elements differ in their bits" % bad
