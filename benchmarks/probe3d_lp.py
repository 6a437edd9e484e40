"""Lane-pair 3D kernels at the NeRF shape (cfg4): forward / backward timings on tile-sorted samples, checked against
the round-1 point-parallel kernels. SHACIRA_LIB selects a tuning build (benchmarks/build_variants.py).
    python benchmarks/probe3d_lp.py [--tp 128] [--tag name]"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shacira_b200 import _lib  # noqa: E402
from shacira_b200.grids import geometric_resolutions  # noqa: E402
from probe3d import setenv, timed  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tp", type=int, default=128)
ap.add_argument("--tag", default="")
ap.add_argument("--bwd", action="store_true")
args = ap.parse_args()
L, BW, C, F, S, SETS = 16, 19, 1, 4, 4096 * 128, 3
dev = torch.device("cuda", 0)
res = geometric_resolutions(16, 2048, L)
sizes = [min(2 ** BW, r ** 3) for r in res]
first = [0]
for s in sizes[:-1]:
    first.append(first[-1] + s)
T = sum(sizes)
torch.manual_seed(7)
lat = (torch.rand((T, C), device=dev) - 0.5) * 16
A = torch.randn((1, C, F), device=dev) * 0.1
shift = torch.randn((1, F), device=dev) * 0.05
sets = [dict(coords=torch.rand((S, 3), device=dev) * 2 - 1, g=torch.randn((S, L * F), device=dev)) for _ in range(SETS)]


def emit(**kw):
    kw["tag"] = args.tag
    print(json.dumps(kw), flush=True)


setenv(SHACIRA_3D_MERGE=0, SHACIRA_3D_RED=-1)
f0, z0 = _lib.latent_forward(sets[0]["coords"], lat, first, res, BW, A, shift, F, True, True)
gl0, gA0, gS0 = _lib.latent_backward(sets[0]["coords"], sets[0]["g"], z0, first, res, BW, A, C, F, T, True)
setenv(SHACIRA_3D_MERGE=None, SHACIRA_3D_RED=None)
plans = [_lib.Plan(s["coords"], tile_points=args.tp) for s in sets]
info = plans[0].info()
perm = plans[0].perm_tensor()
f, z = _lib.latent_forward_planned_z(plans[0], lat, first, res, BW, A, shift, F, True, True)
ok = bool(torch.equal(f, f0) and torch.equal(z, z0[perm]))
fb, zb = torch.empty_like(f), torch.empty_like(z)
us = timed(lambda i: _lib.latent_forward_planned_z(plans[i % SETS], lat, first, res, BW, A, shift, F, True, True, fb, zb), iters=20)
emit(kernel="fwd", variant="lane pairs, sorted g=%d" % info["tiles_per_axis"], us=us, bit_identical=ok)
us = timed(lambda i: _lib.latent_forward(sets[i % SETS]["coords"], lat, first, res, BW, A, shift, F, True, True), iters=20)
emit(kernel="fwd", variant="lane pairs, unsorted", us=us)
if args.bwd:
    zs = [_lib.latent_forward_planned_z(plans[k], lat, first, res, BW, A, shift, F, True, True)[1] for k in range(SETS)]
    for staged, ctas in ((None, None), (None, 3), (5, None), (6, None), (0, None)):
        if True:
            setenv(SHACIRA_3D_STAGED=staged, SHACIRA_3D_BWD_CTAS=ctas)
            for dec in (False, True):
                gl, gA, gS = _lib.latent_backward_planned_z(plans[0], sets[0]["g"], zs[0] if dec else None, first, res, BW, A, C, F, T, dec)
                us = timed(lambda i: _lib.latent_backward_planned_z(plans[i % SETS], sets[i % SETS]["g"], zs[i % SETS] if dec else None, first, res, BW, A, C, F, T, dec), iters=20)
                rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
                lvl = max(float((gl[first[l]:first[l] + sizes[l]] - gl0[first[l]:first[l] + sizes[l]]).abs().max() / gl0[first[l]:first[l] + sizes[l]].abs().max()) for l in range(L))
                emit(kernel="bwd+dec" if dec else "bwd", variant="sorted g=%d staged=%s ctas/sm=%s" % (info["tiles_per_axis"], staged, ctas), us=us,
                     rel=rel(gl, gl0), level_rel=lvl, rel_gA=rel(gA, gA0) if dec else None, rel_gS=rel(gS, gS0) if dec else None)
    setenv(SHACIRA_3D_STAGED=None, SHACIRA_3D_BWD_CTAS=None)
