"""Round-2 probe of the 3D (NeRF-shape, BASELINE cfg4) kernels: every variant of the forward / backward timed on the
same inputs and checked against the round-1 point-parallel kernels. One JSON line per variant.

    python benchmarks/probe3d.py [--quick]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shacira_b200 import _lib  # noqa: E402
from shacira_b200.grids import geometric_resolutions  # noqa: E402

L, BW, C, F = 16, 19, 1, 4
S = 4096 * 128
SETS = 3


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


def timed(fn, iters=12, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
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
    out = []

    def emit(**kw):
        print(json.dumps(kw), flush=True)
        out.append(kw)

    # ---- round-1 kernels: the yardstick ---------------------------------------------------------------------
    setenv(SHACIRA_3D_MERGE=0, SHACIRA_3D_RED=-1)
    ref = []
    for s in sets:
        f, z = _lib.latent_forward(s["coords"], lat, first, res, BW, A, shift, F, True, True)
        gl, gA, gS = _lib.latent_backward(s["coords"], s["g"], z, first, res, BW, A, C, F, T, True)
        ref.append(dict(f=f, z=z, gl=gl, gA=gA, gS=gS))
    us = timed(lambda i: _lib.latent_forward(sets[i % SETS]["coords"], lat, first, res, BW, A, shift, F, True, True))
    emit(kernel="fwd", variant="r01 point-parallel", us=us)
    us = timed(lambda i: _lib.latent_backward(sets[i % SETS]["coords"], sets[i % SETS]["g"], ref[i % SETS]["z"], first, res, BW, A, C, F, T, True))
    emit(kernel="bwd+dec", variant="r01 point-parallel + coarse", us=us)
    us = timed(lambda i: _lib.latent_backward(sets[i % SETS]["coords"], sets[i % SETS]["g"], None, first, res, BW, A, C, F, T, False))
    emit(kernel="bwd", variant="r01 point-parallel + coarse", us=us)

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())

    def level_rel(gl, want):
        worst = 0.0
        for l in range(L):
            a, b = gl[first[l]:first[l] + sizes[l]], want[first[l]:first[l] + sizes[l]]
            worst = max(worst, float((a - b).abs().max() / b.abs().max()))
        return worst

    # ---- unplanned: merged loads / vector reds / lane pairs ------------------------------------------------------
    for merge, name in ((1, "quad-merged x-pairs"), (2, "lane pairs")):
        setenv(SHACIRA_3D_MERGE=merge)
        f, z = _lib.latent_forward(sets[0]["coords"], lat, first, res, BW, A, shift, F, True, True)
        us = timed(lambda i: _lib.latent_forward(sets[i % SETS]["coords"], lat, first, res, BW, A, shift, F, True, True))
        emit(kernel="fwd", variant="%s, unsorted" % name, us=us, bit_identical=bool(torch.equal(f, ref[0]["f"]) and torch.equal(z, ref[0]["z"])))
    for red in (4, 8):
        setenv(SHACIRA_3D_RED=red)
        for dec in (False, True):
            gl, gA, gS = _lib.latent_backward(sets[0]["coords"], sets[0]["g"], ref[0]["z"] if dec else None, first, res, BW, A, C, F, T, dec)
            us = timed(lambda i: _lib.latent_backward(sets[i % SETS]["coords"], sets[i % SETS]["g"], ref[i % SETS]["z"] if dec else None, first, res, BW, A, C, F, T, dec))
            emit(kernel="bwd+dec" if dec else "bwd", variant="unsorted, red mode %d + coarse" % red, us=us,
                 rel=rel(gl, ref[0]["gl"]), level_rel=level_rel(gl, ref[0]["gl"]),
                 rel_gA=rel(gA, ref[0]["gA"]) if dec else None, rel_gS=rel(gS, ref[0]["gS"]) if dec else None)

    # ---- planned (tile-sorted samples) ----------------------------------------------------------------------
    for tp in (384, 128):
        plans = [_lib.Plan(s["coords"], tile_points=tp) for s in sets]
        info = plans[0].info()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        us = timed(lambda i: _lib._check(_lib.load().shacira_plan_rebuild(plans[i % SETS].handle, 3, _lib._ptr(sets[i % SETS]["coords"]), S, tp, st)))
        emit(kernel="plan_rebuild", variant="tile_points=%d" % tp, us=us, tiles=info["ntiles"], g=info["tiles_per_axis"])
        perm = plans[0].perm_tensor()
        for merge, name in ((1, "quad-merged x-pairs"), (2, "lane pairs")):
            setenv(SHACIRA_3D_MERGE=merge)
            f, z = _lib.latent_forward_planned_z(plans[0], lat, first, res, BW, A, shift, F, True, True)
            zs = [_lib.latent_forward_planned_z(plans[k], lat, first, res, BW, A, shift, F, True, True)[1] for k in range(SETS)]
            fb = torch.empty_like(f)
            zb = torch.empty_like(z)
            us = timed(lambda i: _lib.latent_forward_planned_z(plans[i % SETS], lat, first, res, BW, A, shift, F, True, True, fb, zb))
            emit(kernel="fwd", variant="%s, sorted g=%d" % (name, info["tiles_per_axis"]), us=us,
                 bit_identical=bool(torch.equal(f, ref[0]["f"]) and torch.equal(z, ref[0]["z"][perm])))
            us = timed(lambda i: _lib.latent_forward_planned_z(plans[i % SETS], lat, first, res, BW, A, shift, F, True, False, fb, None))
            emit(kernel="fwd(no z)", variant="%s, sorted g=%d" % (name, info["tiles_per_axis"]), us=us)
        for red, ctas in ((4, None), (8, 2), (8, 4), (8, 6)):
            for staged in ((6, None) if red == 4 else (0, 4, 5, 6, 7, None)):
                setenv(SHACIRA_3D_RED=red, SHACIRA_3D_STAGED=staged, SHACIRA_3D_BWD_CTAS=ctas)
                if ctas not in (None, 4) and staged not in (None, 6):
                    continue
                for dec in (False, True):
                    if dec and staged not in (6, None):
                        continue
                    try:
                        gl, gA, gS = _lib.latent_backward_planned_z(plans[0], sets[0]["g"], zs[0] if dec else None, first, res, BW, A, C, F, T, dec)
                        us = timed(lambda i: _lib.latent_backward_planned_z(plans[i % SETS], sets[i % SETS]["g"], zs[i % SETS] if dec else None, first, res, BW, A, C, F, T, dec))
                        emit(kernel="bwd+dec" if dec else "bwd", variant="sorted g=%d, red mode %d, staged=%s, ctas/sm=%s" % (info["tiles_per_axis"], red, staged, ctas),
                             us=us, rel=rel(gl, ref[0]["gl"]), level_rel=level_rel(gl, ref[0]["gl"]),
                             rel_gA=rel(gA, ref[0]["gA"]) if dec else None, rel_gS=rel(gS, ref[0]["gS"]) if dec else None)
                    except Exception as e:
                        emit(kernel="bwd", variant="sorted g=%d, red mode %d, staged=%s" % (info["tiles_per_axis"], red, staged), error=repr(e)[:300])
        setenv(SHACIRA_3D_STAGED=None, SHACIRA_3D_RED=None, SHACIRA_3D_BWD_CTAS=None, SHACIRA_3D_MERGE=None)
        for p in plans:
            p.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r02b_probe3d.jsonl"), "w") as fh:
        for o in out:
            fh.write(json.dumps(o) + "\n")


if __name__ == "__main__":
    main()
