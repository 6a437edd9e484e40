"""Device time of the tiled forward / backward kernels alone, host launch overhead removed by CUDA-graph
replay (K launches per graph over rotating input sets). Used for kernel tuning; bench.py is the contract."""
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


def main():
    K = int(os.environ.get("K", "40"))
    dev = torch.device("cuda", 0)
    wl = bench.make_workload(0)
    d = lambda a: torch.from_numpy(a).to(dev)
    sets = [dict(coords=d(s["coords"]), grad_out=d(s["grad_out"])) for s in wl["sets"]]
    lat, A, shift = d(wl["latents"]), d(wl["A"]), d(wl["shift"])
    first, res, T, L = wl["first"], wl["res"], wl["T"], bench.NUM_LODS
    n = bench.H * bench.W
    lib = _lib.load()
    fi, _ = _lib._i32_array(first)
    rs, _ = _lib._i32_array(res)
    for s in sets:
        s["plan"] = _lib.Plan(s["coords"])
        s["feats"] = torch.empty((n, L), device=dev)
        s["gl"] = torch.empty((T, 1), device=dev)
        s["z"] = torch.empty((n, L), device=dev)
    gA = torch.zeros((L, 1, 1), device=dev)
    gS = torch.zeros((L, 1), device=dev)
    p = _lib._ptr

    def fwd(s, st):
        _lib._check(lib.shacira_latent_forward_planned(s["plan"].handle, p(lat), fi, rs, L, bench.BITWIDTH, 1, 1, 1, p(A),
                                                       p(shift), 0, p(s["feats"]), st))

    def bwd(s, st, dec=True):
        _lib._check(lib.shacira_latent_backward_planned(s["plan"].handle, p(s["grad_out"]), p(lat), fi, rs, L,
                                                        bench.BITWIDTH, 1, 1, 1, p(A), 0, T, 1, p(s["gl"]),
                                                        p(gA) if dec else None, p(gS) if dec else None, st))

    for s_ in sets:
        s_["gmax"] = s_["grad_out"].abs().amax(dim=0).contiguous()

    def bwd_bounded(s, st):
        _lib._check(lib.shacira_latent_backward_planned_bounded(s["plan"].handle, p(s["grad_out"]), p(lat), fi, rs, L,
                                                                bench.BITWIDTH, 1, 1, 1, p(A), 0, T, 1, p(s["gl"]),
                                                                p(gA), p(gS), p(s["gmax"]), st))

    for s_ in sets:
        s_["plan_sorted"] = _lib.Plan(s_["coords"]).set_sorted_io(True)

    def fwd_sorted(s, st):
        _lib._check(lib.shacira_latent_forward_planned(s["plan_sorted"].handle, p(lat), fi, rs, L, bench.BITWIDTH, 1, 1, 1,
                                                       p(A), p(shift), 0, p(s["feats"]), st))

    def bwd_sorted_bounded(s, st):
        _lib._check(lib.shacira_latent_backward_planned_bounded(s["plan_sorted"].handle, p(s["grad_out"]), p(lat), fi, rs,
                                                                L, bench.BITWIDTH, 1, 1, 1, p(A), 0, T, 1, p(s["gl"]),
                                                                p(gA), p(gS), p(s["gmax"]), st))

    def fwd_pp(s, st):
        _lib._check(lib.shacira_latent_forward(2, p(s["coords"]), n, p(lat), fi, rs, L, bench.BITWIDTH, 1, 1, 1, p(A),
                                               p(shift), 0, p(s["feats"]), p(s["z"]), st))

    def bwd_pp(s, st):
        _lib._check(lib.shacira_latent_backward(2, p(s["coords"]), n, p(s["grad_out"]), p(s["z"]), fi, rs, L,
                                                bench.BITWIDTH, 1, 1, p(A), 0, T, 1, p(s["gl"]), p(gA), p(gS), st))

    noise, prob = d(wl["noise"]), d(wl["prob"])
    bits = torch.empty((1 + L,), dtype=torch.float64, device=dev)
    egl = torch.empty((T, 1), device=dev)
    egp = torch.empty((4, 3, 1), device=dev)

    scratch = torch.zeros(int(lib.shacira_entropy_scratch_bytes(1, L)), dtype=torch.uint8, device=dev)

    def ent(s, st):
        _lib._check(lib.shacira_entropy_bits(p(lat), p(noise), T, 1, p(prob), 2, fi, L, p(bits), p(egl), p(egp),
                                             p(scratch), scratch.numel(), st))

    import torch.nn as nn
    torch.manual_seed(0)
    mlp = nn.Sequential(nn.Linear(16, 16), nn.ReLU(), nn.Linear(16, 16), nn.ReLU(), nn.Linear(16, 3)).to(dev)
    gt = torch.rand(n, 3, device=dev)
    mgx = torch.empty((n, L), device=dev)
    mout = torch.empty(2 + 16 * 16 + 16 + 256 + 16 + 48 + 3, device=dev)
    mw = [mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias, mlp[4].weight, mlp[4].bias]

    def mlpk(s, st):
        _lib._check(lib.shacira_mlp_mse_step(p(s["feats"]), p(gt), n, 16, 16, 3, *[p(w) for w in mw], p(mgx), None, p(mout), st))

    gmaxbuf = torch.zeros(16, device=dev)
    mout2 = torch.zeros(2 + 16 * 16 + 16 + 256 + 16 + 48 + 3 + 16, device=dev)
    gl_acc = torch.zeros((T, 1), device=dev)

    def three_sorted(s, st):   # what the fit step ran before the fused kernel: rows in tile order, bound from the MLP
        _lib._check(lib.shacira_latent_forward_planned(s["plan_sorted"].handle, p(lat), fi, rs, L, bench.BITWIDTH, 1, 1, 1,
                                                       p(A), p(shift), 0, p(s["feats"]), st))
        _lib._check(lib.shacira_mlp_mse_step_bounded(p(s["feats"]), p(gt), n, 16, 16, 3, *[p(w) for w in mw], p(mgx), None,
                                                     p(mout2), p(mout2[-16:]), st))
        _lib._check(lib.shacira_latent_backward_planned_bounded(s["plan_sorted"].handle, p(mgx), p(lat), fi, rs, L,
                                                                bench.BITWIDTH, 1, 1, 1, p(A), 0, T, 0, p(gl_acc), p(gA),
                                                                p(gS), p(mout2[-16:]), st))

    def fused(s, st):
        _lib._check(lib.shacira_fit_tile_step(s["plan_sorted"].handle, p(lat), fi, rs, L, bench.BITWIDTH, 1, p(A), p(shift),
                                              p(gt), *[p(w) for w in mw], T, p(gl_acc), p(gA), p(gS), p(mout2), st))

    out = {}
    only_ent = bool(os.environ.get("ONLY_ENT"))
    for name, fn in ((("entropy_fwd_bwd", ent),) if only_ent else (("entropy_fwd_bwd", ent), ("mlp_mse_step", mlpk), ("fwd_tiled", fwd), ("bwd_tiled_dec", bwd), ("bwd_tiled_nodec", lambda s, st: bwd(s, st, False)), ("bwd_tiled_dec_bounded", bwd_bounded), ("fwd_tiled_sorted_io", fwd_sorted),
                     ("bwd_tiled_dec_bounded_sorted_io", bwd_sorted_bounded),
                     ("fwd_pointparallel", fwd_pp), ("bwd_pointparallel", bwd_pp),
                     ("step_tiled", lambda s, st: (fwd(s, st), bwd(s, st))),
                     ("fit_three_kernels_sorted_io", three_sorted), ("fit_tile_fused", fused))):
        if os.environ.get("ONLY") and name not in os.environ["ONLY"].split(","):
            continue
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            st = ctypes.c_void_p(stream.cuda_stream)
            for i in range(4):
                fn(sets[i % len(sets)], st)
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                st2 = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                for i in range(K):
                    fn(sets[i % len(sets)], st2)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(5):
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / K * 1e3)
        out[name] = round(best, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
