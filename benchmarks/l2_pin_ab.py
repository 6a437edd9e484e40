"""A/B of the L2 access-policy window (shacira_l2_pin, north_star: "pinned in the 126 MB L2 via access-policy windows") on
the tables of the two headline shapes: NeRF 3D grid (24.4 MB table beside 134 MB gradient / feature streams per step) and
the Kodak 2D grid (1.5 MB table). Device time per forward + backward (CUDA events, eager launches on one stream)."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
from shacira_b200 import _lib  # noqa: E402
from shacira_b200.grids import geometric_resolutions  # noqa: E402


def run3d(dev, pin, steps=30):
    L, BW, C, F = 16, 19, 1, 4
    res = geometric_resolutions(16, 2048, L)
    sizes = [min(2 ** BW, r ** 3) for r in res]
    first = [0]
    for s in sizes[:-1]:
        first.append(first[-1] + s)
    T, S = sum(sizes), 4096 * 128
    torch.manual_seed(0)
    lat = (torch.rand((T, C), device=dev) - 0.5) * 16
    A = torch.randn((1, C, F), device=dev) * 0.1
    shift = torch.zeros((1, F), device=dev)
    sets = [dict(c=torch.rand((S, 3), device=dev) * 2 - 1, g=torch.randn((S, L * F), device=dev)) for _ in range(3)]
    feats, z = torch.empty((S, L * F), device=dev), torch.empty((S, L * C), device=dev)
    gl, gA, gS = torch.zeros((T, C), device=dev), torch.zeros((L, C, F), device=dev), torch.zeros((L, F), device=dev)
    lib, P = _lib.load(), _lib._ptr
    fi, _ = _lib._i32_array(first)
    rs, _ = _lib._i32_array(res)
    plans = [_lib.Plan(s["c"]) for s in sets]
    st = torch.cuda.Stream(device=dev)
    out = {}
    with torch.cuda.stream(st):
        if pin:
            _lib.l2_pin(lat)
        h = ctypes.c_void_p(st.cuda_stream)

        def step(i):
            s, p = sets[i % 3], plans[i % 3]
            _lib._check(lib.shacira_latent_forward_planned_z(p.handle, P(lat), fi, rs, L, BW, C, F, 1, P(A), P(shift), 0,
                                                             P(feats), P(z), h))
            _lib._check(lib.shacira_latent_backward_planned_z(p.handle, P(s["g"]), P(z), fi, rs, L, BW, C, F, P(A), 0, T, 0,
                                                              P(gl), P(gA), P(gS), h))
        for i in range(5):
            step(i)
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        st.synchronize()
        out = e0.elapsed_time(e1) / steps * 1e3
        if pin:
            _lib.l2_pin(None)
    for p in plans:
        p.close()
    return out


def main():
    dev = torch.device("cuda", 0)
    info = _lib.device_info() if hasattr(_lib, "device_info") else {}
    r = {"device": info}
    for rep in range(2):
        r["nerf3d_fwd_bwd_us_unpinned_%d" % rep] = run3d(dev, False)
        r["nerf3d_fwd_bwd_us_pinned_%d" % rep] = run3d(dev, True)
    print(json.dumps(r))


if __name__ == "__main__":
    main()
