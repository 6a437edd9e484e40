"""Probe for the next round: do the 3D point-parallel kernels gain from spatially sorted samples? Same kernels, same
table, NeRF shape; coordinates either random (as nerf_dp.py draws them) or in the tile order of a 3D plan."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shacira_b200 import _lib  # noqa: E402
from shacira_b200.grids import geometric_resolutions  # noqa: E402
from kernel_times_3d import timed  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    L, BW, S, C, F = 16, 19, 1 << 19, 1, 4
    res = geometric_resolutions(16, 2048, L)
    sizes = [min(2 ** BW, r ** 3) for r in res]
    first = [0]
    for s in sizes[:-1]:
        first.append(first[-1] + s)
    T = sum(sizes)
    lib = _lib.load()
    fi, _ = _lib._i32_array(first)
    rs, _ = _lib._i32_array(res)
    P = _lib._ptr
    torch.manual_seed(0)
    lat = (torch.rand((T, C), device=dev) - 0.5) * 16
    A = torch.randn((1, C, F), device=dev) * 0.1
    shift = torch.zeros((1, F), device=dev)
    feats = torch.empty((S, L * F), device=dev)
    z = torch.empty((S, L * C), device=dev)
    gl = torch.zeros((T, C), device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    sets = []
    for k in range(4):
        c = torch.rand((S, 3), device=dev) * 2 - 1
        plan = _lib.Plan(c)
        cs = torch.from_numpy(plan.arrays()[1]).to(dev).contiguous()
        plan.close()
        sets.append((c, cs, torch.randn((S, L * F), device=dev)))
    for name, idx in (("random", 0), ("tile_sorted", 1)):
        fwd = lambda i: _lib._check(lib.shacira_latent_forward(3, P(sets[i % 4][idx]), S, P(lat), fi, rs, L, BW, C, F, 1, P(A),
                                                               P(shift), 0, P(feats), P(z), st))
        bwd = lambda i: _lib._check(lib.shacira_latent_backward(3, P(sets[i % 4][idx]), S, P(sets[i % 4][2]), None, fi, rs, L,
                                                                BW, C, F, P(A), 0, T, 1, P(gl), None, None, st))
        out[name] = {"fwd_us": timed(fwd), "bwd_us": timed(bwd)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
