"""Capture outputs of the REFERENCE'S OWN CUDA kernels (oracle/_ref, compiled unmodified from
/root/reference/wisp/csrc/ops for sm_100a) on a B200 into tests/golden/hashgrid_ref_kernels.npz.

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/hashgrid_ref_kernels.npz'

The committed fixture lets the CPU-only test suite pin the C oracle against what the reference kernels
really produce (forward: bit for bit; backward: float atomics, order-dependent, so to 1e-5)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from helpers import make_case  # noqa: E402
from oracle import build_ref  # noqa: E402

CASES = {  # name: dim, L, bw, rmin, rmax, n, F, coords
    "2d_cfg1": (2, 16, 14, 16, 512, 600, 2, "uniform"),
    "2d_cfg2_arbitrary": (2, 16, 16, 16, 512, 600, 2, "arbitrary"),
    "2d_q4_dense": (2, 8, 19, 16, 700, 400, 2, "uniform"),
    "3d_cfg4": (3, 16, 19, 16, 2048, 600, 2, "arbitrary"),
    "3d_lego24_f4": (3, 24, 19, 16, 512, 300, 4, "uniform"),
}


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "hashgrid_ref_kernels.npz")
    ref = build_ref.load()
    assert ref is not None and torch.cuda.is_available()
    out = {"cases": np.array(sorted(CASES))}
    for name, (dim, L, bw, rmin, rmax, n, F, kind) in CASES.items():
        c = make_case(dim, L, bw, rmin, rmax, n, F, seed=len(name) * 7 + dim, coord_kind=kind)
        edge = np.array([[-1.0] * dim, [1.0] * dim, [0.0] * dim, [0.9999999] * dim, [1.25] * dim, [-3.0] * dim], np.float32)
        c["coords"][:len(edge)] = edge
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        first = torch.tensor(c["first_idx"], dtype=torch.int32, device="cuda")
        fwd = ref.hashgrid_interpolate2d_cuda if dim == 2 else ref.hashgrid_interpolate_cuda
        bwd = ref.hashgrid_interpolate2d_backward_cuda if dim == 2 else ref.hashgrid_interpolate_backward_cuda
        feats = fwd(dev(c["coords"]), dev(c["table"]), first, c["resolutions"], bw)
        grad = bwd(dev(c["coords"]), dev(c["grad_out"]), dev(c["table"]), first, c["resolutions"], bw, F, False)
        p = name + "/"
        out[p + "meta"] = np.array([dim, L, bw, F], np.int32)
        out[p + "resolutions"] = np.array(c["resolutions"], np.int32)
        out[p + "coords"] = c["coords"]
        out[p + "table_seed_check"] = c["table"][:16].copy()   # the table is regenerated from the seed (kept small)
        out[p + "seed"] = np.array([len(name) * 7 + dim], np.int64)
        out[p + "feats"] = feats.cpu().numpy()
        # gradients: a fixed sample of the nonzero rows + the per-column totals (small fixture)
        g = grad.cpu().numpy()
        nz = np.nonzero(np.abs(g).sum(1))[0].astype(np.int32)
        pick = nz[np.random.default_rng(0).permutation(len(nz))[:1500]]
        out[p + "grad_rows"] = np.sort(pick)
        out[p + "grad_vals"] = g[np.sort(pick)]
        out[p + "grad_nonzero_rows"] = np.array([len(nz)], np.int64)
        out[p + "grad_total"] = g.astype(np.float64).sum(0)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main()
