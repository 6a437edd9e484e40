import sys, torch
sys.path.insert(0, "benchmarks")
import fit_image
dev = torch.device("cuda", 0)
for seed in (0, 1, 2):
    for impl in ("ours", "ref"):
        r = fit_image.fit(seed, impl, 600, dev, use_graph=False, noise_cpu=True)
        print(seed, impl, round(r["psnr"], 3), round(r["bpp"], 4), round(r["rgb_loss"], 5), flush=True)
