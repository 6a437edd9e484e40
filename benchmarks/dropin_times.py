"""Device time of the reference's own entry points (wisp._C.ops names) at the Kodak shape as the reference calls them:
F = 2 after LatentGrid.interpolate's repeat(1, 2) (latent_grid.py:361-363). Tiled path (plan cached per coordinate
tensor) vs SHACIRA_DISABLE_PLAN=1 (point-parallel kernels). CUDA events, 50 iterations after 5 warm-ups."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from shacira_b200._C import ops  # noqa: E402


def timed(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    dev = torch.device("cuda", 0)
    wl = bench.make_workload(0)
    coords = torch.from_numpy(wl["sets"][0]["coords"]).to(dev)
    n, L, T = coords.shape[0], bench.NUM_LODS, wl["T"]
    table = torch.randn((T, 2), device=dev)
    g = torch.randn((n, 2 * L), device=dev)
    first = torch.tensor(wl["first"], dtype=torch.int32, device=dev)
    res = wl["res"]
    fwd = lambda: ops.hashgrid_interpolate2d_cuda(coords, table, first, res, bench.BITWIDTH)
    bwd = lambda: ops.hashgrid_interpolate2d_backward_cuda(coords, g, table, first, res, bench.BITWIDTH, 2, False)
    print(json.dumps({"shape": "Kodak 768x512, 16 levels, 2^16 rows, F=2 (reference call pattern)",
                      "plan": "disabled" if os.environ.get("SHACIRA_DISABLE_PLAN") else "cached per coords tensor",
                      "fwd_us": timed(fwd), "bwd_us": timed(bwd)}))


if __name__ == "__main__":
    main()
