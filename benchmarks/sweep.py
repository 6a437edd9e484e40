"""Hash-grid microbench sweep (BASELINE cfg5 / cfg4 shapes): table 2^14..2^22 rows x 2^16..2^22 points, plain HashGrid
(F=2) and LatentGrid (C=1 -> F=4), 3D (and the 2D image grid), forward + backward device time by CUDA-graph
replay. Honours the reference's int32-overflow window: max resolution capped at 1290 for 2^21+ rows.
    python benchmarks/sweep.py [--quick]
One JSON line per configuration: Mpoints/s fwd+bwd, algorithmic GB/s and fraction of the measured HBM peak."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from shacira_b200 import _lib  # noqa: E402
from shacira_b200.grids import geometric_resolutions  # noqa: E402


def layout(res, bw, dim):
    sizes = [min(2 ** bw, r ** dim) for r in res]
    first = [0]
    for s in sizes[:-1]:
        first.append(first[-1] + s)
    return first, sum(sizes)


def time_graph(fn, K=10):
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        st = ctypes.c_void_p(stream.cuda_stream)
        for _ in range(2):
            fn(st)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            st2 = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            for _ in range(K):
                fn(st2)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / K)
    return best


def run(dim, L, bw, rmin, rmax, n, C, F, planned, peak):
    dev = torch.device("cuda", 0)
    res = geometric_resolutions(rmin, rmax, L)
    first, T = layout(res, bw, dim)
    torch.manual_seed(0)
    coords = torch.rand((n, dim), device=dev) * 2 - 1
    lat = (torch.rand((T, C), device=dev) - 0.5) * 16
    A = torch.randn((1, C, F), device=dev) * 0.1
    shift = torch.zeros((1, F), device=dev)
    g = torch.randn((n, L * F), device=dev)
    feats = torch.empty((n, L * F), device=dev)
    z = torch.empty((n, L * C), device=dev)
    gl = torch.empty((T, C), device=dev)
    gA = torch.zeros((L, C, F), device=dev)
    gS = torch.zeros((L, F), device=dev)
    lib = _lib.load()
    fi, _ = _lib._i32_array(first)
    rs, _ = _lib._i32_array(res)
    P = _lib._ptr
    plan = _lib.Plan(coords) if planned else None

    def fwd(st):
        if planned:
            _lib._check(lib.shacira_latent_forward_planned(plan.handle, P(lat), fi, rs, L, bw, C, F, 1, P(A), P(shift), 0,
                                                           P(feats), st))
        else:
            _lib._check(lib.shacira_latent_forward(dim, P(coords), n, P(lat), fi, rs, L, bw, C, F, 1, P(A), P(shift), 0,
                                                   P(feats), P(z), st))

    def bwd(st):
        if planned:
            _lib._check(lib.shacira_latent_backward_planned(plan.handle, P(g), P(lat), fi, rs, L, bw, C, F, 1, P(A), 0, T, 1,
                                                            P(gl), P(gA), P(gS), st))
        else:
            _lib._check(lib.shacira_latent_backward(dim, P(coords), n, P(g), P(z), fi, rs, L, bw, C, F, P(A), 0, T, 1,
                                                    P(gl), P(gA), P(gS), st))

    def replan(st):
        _lib._check(lib.shacira_plan_rebuild(plan.handle, dim, P(coords), n, 0, st))

    tf, tb = time_graph(fwd), time_graph(bwd)
    tp = time_graph(replan) if planned else 0.0
    bf, bb = bench.algorithmic_bytes_per_point(dim, L, C, F)
    out = {"dim": dim, "levels": L, "log2_rows": bw, "max_res": rmax, "points": n, "C": C, "F": F,
           "path": "tiled" if planned else "point-parallel", "table_MB": round(T * C * 4 / 2 ** 20, 1),
           "fwd_us": round(tf * 1e3, 1), "bwd_us": round(tb * 1e3, 1), "replan_us": round(tp * 1e3, 1),
           "Mpts_s": round(n / (tf + tb) / 1e3, 1), "alg_GBs": round((bf + bb) * n / (tf + tb) / 1e6, 1),
           "frac_hbm_peak": round((bf + bb) * n / (tf + tb) / 1e6 / peak, 3)}
    print(json.dumps(out), flush=True)
    if plan is not None:
        plan.close()


def main():
    quick = "--quick" in sys.argv
    peak, _ = bench.measured_peaks()
    # BASELINE cfg2 / cfg4 shapes first
    for planned in (False, True):
        run(2, 16, 16, 16, 512, 768 * 512, 1, 1, planned, peak)
        run(3, 16, 19, 16, 2048, 1 << 19, 1, 4, planned, peak)
    if quick:
        return
    for bw in (14, 16, 18, 20, 22):
        rmax = 2048 if bw < 21 else 1290  # SURVEY Q2 window
        for logn in (16, 18, 20, 22):
            for (C, F) in ((2, 2), (1, 4)):
                # 3D runs on the point-parallel kernels (+ coarse shared-memory backward); --tiled3d adds the
                # tiled path, measured slower there (most 3D levels do not fit a tile's shared-memory box)
                for planned in ((False, True) if "--tiled3d" in sys.argv else (False,)):
                    run(3, 16, bw, 16, rmax, 1 << logn, C, F, planned, peak)


if __name__ == "__main__":
    main()
