import ctypes, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from shacira_b200 import _lib
dev = torch.device("cuda", 0)
wl = bench.make_workload(0)
d = lambda a: torch.from_numpy(a).to(dev)
lat, noise, prob = d(wl["latents"]), d(wl["noise"]), d(wl["prob"])
for _ in range(5):
    _lib.entropy_bits(lat, noise, prob, 2, wl["first"])
torch.cuda.synchronize()
