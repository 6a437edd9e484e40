"""NeRF-shape ray-batch data parallel step (BASELINE cfg4): 3D LatentGrid, 16 levels 16->2048, 2^19-row tables,
C=1 -> F=4; 4096 rays x 128 samples (8 cells x 16 steps, synthetic sampler: kaolin's raymarcher is out of scope)
per rank and step; forward + backward of the grid, then the ONE exchange step of the path: NCCL SUM all-reduce of
grad(latents) (24 MB fp32) plus a flat bucket with the decoder gradients.

    python benchmarks/nerf_dp.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/nerf_dp.py

Weak scaling (4096 rays per rank). Device timing (CUDA events), barrier on both sides, max over ranks."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from shacira_b200 import _lib, dp  # noqa: E402
from shacira_b200.grids import geometric_resolutions  # noqa: E402

RAYS, SAMPLES_PER_RAY = 4096, 128
L, BW, C, F = 16, 19, 1, 4


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps, warmup, sets_n = 30, 5, 4
    chunks = 1
    if "--chunks" in sys.argv:
        chunks = int(sys.argv[sys.argv.index("--chunks") + 1])
    res = geometric_resolutions(16, 2048, L)
    sizes = [min(2 ** BW, r ** 3) for r in res]
    first = [0]
    for s in sizes[:-1]:
        first.append(first[-1] + s)
    T = sum(sizes)
    S = RAYS * SAMPLES_PER_RAY
    torch.manual_seed(1234)  # replicated parameters
    lat = (torch.rand((T, C), device=dev) - 0.5) * 16
    A = torch.randn((1, C, F), device=dev) * 0.1
    shift = torch.zeros((1, F), device=dev)
    torch.manual_seed(100 + rank)  # this rank's rays
    sets = [dict(coords=torch.rand((S, 3), device=dev) * 2 - 1, g=torch.randn((S, L * F), device=dev)) for _ in range(sets_n)]
    feats = torch.empty((S, L * F), device=dev)
    z = torch.empty((S, L * C), device=dev)
    glat = torch.nn.Parameter(torch.zeros((T, C), device=dev))
    gA = torch.nn.Parameter(torch.zeros((L, C, F), device=dev))
    gS = torch.nn.Parameter(torch.zeros((L, F), device=dev))
    arena = dp.GradArena([glat, gA, gS])     # one flat gradient buffer: the exchange step is ONE all-reduce
    lib = _lib.load()
    fi, _ = _lib._i32_array(first)
    rs, _ = _lib._i32_array(res)
    P = _lib._ptr

    # level chunks with about equal numbers of rows: [level_lo, level_hi), flat range of the rows in the arena
    bounds = [0]
    for k in range(1, chunks):
        tgt = T * k // chunks
        l = min(range(1, L), key=lambda l_: abs(first[l_] - tgt))
        bounds.append(max(l, bounds[-1] + 1))
    bounds.append(L)
    spans = []
    for k in range(chunks):
        lo, hi = bounds[k], bounds[k + 1]
        mask = sum(1 << l for l in range(lo, hi))
        r0 = first[lo] * C
        r1 = (first[hi] * C) if hi < L else arena.flat.numel()      # the last chunk carries the decoder gradients too
        spans.append((mask, r0, r1))

    def step_chunked(i):
        """Backward in level chunks; the all-reduce of a chunk's rows is in flight while the next chunk computes."""
        s = sets[i % sets_n]
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib._check(lib.shacira_latent_forward(3, P(s["coords"]), S, P(lat), fi, rs, L, BW, C, F, 1, P(A), P(shift), 0,
                                               P(feats), P(z), st))
        gA.grad.zero_()
        gS.grad.zero_()
        handles = []
        for k, (mask, r0, r1) in enumerate(spans):
            _lib._check(lib.shacira_latent_backward_levels(3, P(s["coords"]), S, P(s["g"]), P(z), fi, rs, L, BW, C, F, P(A),
                                                           0, T, 1 if k == 0 else 0, mask, P(glat.grad), P(gA.grad),
                                                           P(gS.grad), st))
            if world > 1:
                handles.append(dist.all_reduce(arena.flat[r0:r1], op=dist.ReduceOp.SUM, async_op=True))
        for h in handles:
            h.wait()
        return len(handles)

    planned = "--unplanned" not in sys.argv
    plan = _lib.Plan(sets[0]["coords"]) if planned else None

    def step(i):
        if chunks > 1:
            return step_chunked(i)
        s = sets[i % sets_n]
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if planned:
            # the samples are new every step: re-bin them (3 small kernels, allocation reused), then the sorted kernels
            _lib._check(lib.shacira_plan_rebuild(plan.handle, 3, P(s["coords"]), S, 0, st))
            _lib._check(lib.shacira_latent_forward_planned_z(plan.handle, P(lat), fi, rs, L, BW, C, F, 1, P(A), P(shift), 0,
                                                             P(feats), P(z), st))
            arena.zero_()   # one memset: table gradient + decoder gradients
            _lib._check(lib.shacira_latent_backward_planned_z(plan.handle, P(s["g"]), P(z), fi, rs, L, BW, C, F, P(A), 0,
                                                              T, 0, P(glat.grad), P(gA.grad), P(gS.grad), st))
            return arena.allreduce()
        _lib._check(lib.shacira_latent_forward(3, P(s["coords"]), S, P(lat), fi, rs, L, BW, C, F, 1, P(A), P(shift), 0,
                                               P(feats), P(z), st))
        gA.grad.zero_()
        gS.grad.zero_()
        _lib._check(lib.shacira_latent_backward(3, P(s["coords"]), S, P(s["g"]), P(z), fi, rs, L, BW, C, F, P(A), 0, T, 1,
                                                P(glat.grad), P(gA.grad), P(gS.grad), st))
        return arena.allreduce()

    for i in range(warmup):
        ncoll = step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item()) / steps
    if rank == 0:
        bf, bb = bench.algorithmic_bytes_per_point(3, L, C, F)
        peak, _ = bench.measured_peaks()
        print(json.dumps({"workload": "BASELINE cfg4 NeRF-shape ray-batch DP step (grid fwd+bwd + grad all-reduce)",
                          "n_gpus": world, "scaling": "weak", "rays_per_rank": RAYS, "samples_per_rank": S,
                          "ms_per_step": ms, "samples_per_s": S * world / ms * 1e3, "rays_per_s": RAYS * world / ms * 1e3,
                          "allreduce_bytes": T * C * 4, "collectives_per_step": ncoll,
                          "alg_GBs_per_gpu": (bf + bb) * S / ms / 1e6, "frac_hbm_peak": (bf + bb) * S / ms / 1e6 / peak,
                          "backward_level_chunks": chunks,
                          "path": "plan re-binned per step + sorted lane-pair kernels + tile-staged coarse levels" if planned
                          else "unsorted lane-pair kernels + whole-level shared-memory coarse path"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
