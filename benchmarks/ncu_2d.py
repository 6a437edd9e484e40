"""A few launches of the 2D (Kodak-shape, cfg2) kernels for an ncu capture:
    ncu --set full ... python benchmarks/ncu_2d.py [--reps 2] [--entropy]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from shacira_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--entropy", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
wl = bench.make_workload(0)
d = lambda a: torch.from_numpy(a).to(dev)
s = wl["sets"][0]
coords, g = d(s["coords"]), d(s["grad_out"])
lat, A, S = d(wl["latents"]), d(wl["A"]), d(wl["shift"])
plan = _lib.Plan(coords)
for _ in range(args.reps):
    f = _lib.latent_forward_planned(plan, lat, wl["first"], wl["res"], bench.BITWIDTH, A, S, 1, True)
    _lib.latent_backward_planned(plan, g, lat, wl["first"], wl["res"], bench.BITWIDTH, A, 1, 1, wl["T"], True, True)
    if args.entropy:
        _lib.entropy_bits(lat, d(wl["noise"]), d(wl["prob"]), 2, wl["first"])
torch.cuda.synchronize()
