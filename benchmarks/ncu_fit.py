"""A few steps of the natively fused Kodak-shape fit (ImageFitStep, SGA on, in-kernel noise) for an ncu capture:
    ncu --set full -k regex:fit_ ... python benchmarks/ncu_fit.py [--steps 4]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
import fit_image  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--ste", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
grid, mlp, coords, gt, fs = fit_image._native_setup(0, dev, device_noise=True, sga=not args.ste)
fs.set_lambda(5e-4)
fs.set_temperature(0.5)
for _ in range(args.steps):
    fs.step()
torch.cuda.synchronize()
print(float(fs.rgb_loss()), float(fs.total_bits()))
