"""Generate tests/golden/sampling_ref.npz by running the REFERENCE's own in-repo sampling helpers
(wisp/ops/spc/sampling.py:35-71: sample_from_depth_intervals, expand_pack_boundary) on seeded inputs. The module is
loaded straight from its file (it only needs torch), so nothing of kaolin is touched. Build container only; the
fixture is committed. The jitter the reference draws with torch.rand_like is reproduced from the same seed."""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_sampling", "/root/reference/wisp/ops/spc/sampling.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
for name, M, K, seed in (("a", 37, 16, 0), ("b", 200, 4, 1), ("c", 5, 1, 2), ("d", 64, 33, 3)):
    g = torch.Generator().manual_seed(100 + seed)
    d0 = torch.rand(M, 1, generator=g) * 3.0
    depth = torch.cat((d0, d0 + torch.rand(M, 1, generator=g) * 0.2 + 1e-3), dim=1)
    torch.manual_seed(seed)
    samples = ref.sample_from_depth_intervals(depth, K)                      # draws rand_like(steps) internally
    torch.manual_seed(seed)
    jitter = torch.rand_like(torch.zeros(M, K))                              # the same draw
    ridx = torch.sort(torch.randint(0, max(2, M // 3), (M,), generator=g))[0].int()
    first = torch.ones(M, dtype=torch.bool)
    first[1:] = ridx[1:] != ridx[:-1]                                        # what spc_render.mark_first_hit returns
    big = ref.expand_pack_boundary(first, K)
    out.update({name + "/depth": depth.numpy(), name + "/jitter": jitter.numpy(), name + "/K": np.array(K),
                name + "/depth_samples": samples.numpy(), name + "/ridx": ridx.numpy(),
                name + "/boundary": big.numpy().astype(np.uint8)})
np.savez_compressed(os.path.join(HERE, "sampling_ref.npz"), **out)
print("wrote", len(out), "arrays")
