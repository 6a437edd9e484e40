"""NeRF-shape ray-batch data parallel step (BASELINE cfg4): 3D LatentGrid, 16 levels 16->2048, 2^19-row tables,
C=1 -> F=4; 4096 rays x 128 samples (8 cells x 16 steps, synthetic sampler: kaolin's raymarcher is out of scope)
per rank and step. One step = re-bin this step's samples (plan rebuild) -> forward -> backward (latents + decoder
gradients), then the ONE exchange step of the path: NCCL SUM all-reduce of one flat gradient arena (grad(latents)
24 MB fp32 + the decoder gradients).

    python benchmarks/nerf_dp.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/nerf_dp.py

Weak scaling (4096 rays per rank). Device timing (CUDA events), barrier on both sides, max over ranks. The samples of
step i+1 do not depend on step i's parameters, so their binning runs on a side stream while step i's gradients are
exchanged (two plans, double buffered): the all-reduce hides it. `run()` is what bench.py calls."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shacira_b200 import _lib, dp, peer  # noqa: E402
from shacira_b200.grids import geometric_resolutions  # noqa: E402

RAYS, SAMPLES_PER_RAY = 4096, 128
L, BW, C, F = 16, 19, 1, 4


def algorithmic_bytes_per_sample():
    """fwd [4D + 2^D L C 4 + 4 L F] + bwd [same] for D = 3."""
    one = 4 * 3 + 8 * L * C * 4 + 4 * L * F
    return 2 * one


def run(dev, rank, world, steps=30, warmup=5, planned=True, overlap_binning=True, sets_n=4, entropy=True):
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
    # bit-rate loss (multiview_trainer.py:110 calls ent_loss with is_val = pipeline.training, SURVEY Q8: the NeRF recipe
    # evaluates it on round(w), so only the density model receives a gradient). It is a sum over table rows: every rank
    # takes a slice (dp.shard_rows), its bits and density-model gradients ride in the same arena and the all-reduce adds
    # the slices -- no second collective, and 1/N of the table-side work per rank.
    gprob = torch.nn.Parameter(torch.zeros((4, 3, C), device=dev))
    gbits = torch.nn.Parameter(torch.zeros((1,), device=dev))
    prob = torch.randn((4, 3, C), device=dev) * 0.3
    r0, r1 = dp.shard_rows(T, rank, world)
    ent_stream = torch.cuda.Stream(device=dev)
    # one flat gradient buffer: the exchange step is ONE launch -- the peer-memory kernel over NVLink (csrc/peer_kernels.cuh,
    # every rank maps every arena through CUDA IPC) or, with SHACIRA_DP_EXCHANGE=nccl / where IPC is unavailable, ncclAllReduce
    arena, exchange_kind = None, "nccl all-reduce"
    # auto: the peer-memory kernel, else nccl. Measured at 8 GPUs (profiles/r02q_peer_exchange.md): peer kernel 91 us,
    # multicast kernel 89 us, ncclAllReduce 124 us for the 24.4 MB arena; the multicast form needs torch's private
    # symmetric-memory allocator for its mapping and is only taken when asked for.
    want = os.environ.get("SHACIRA_DP_EXCHANGE", "auto")
    grads = [glat, gA, gS, gprob, gbits]
    if world in (2, 4, 8) and want == "multimem":
        arena = peer.McArena.try_create(grads)
        if arena is not None:
            exchange_kind = "NVSwitch multicast kernel (one launch per rank: barrier, multimem.ld_reduce own slice, multimem.st to all, barrier)"
    if arena is None and world in (2, 4, 8) and want in ("auto", "peer", "multimem"):
        arena = peer.PeerArena.try_create(grads)
        if arena is not None:
            exchange_kind = "peer-memory kernel over NVLink (one launch per rank: barrier, reduce own slice from all arenas, store to all arenas, barrier)"
    if arena is None:
        arena = dp.GradArena([glat, gA, gS, gprob, gbits])
    lib = _lib.load()
    fi, _ = _lib._i32_array(first)
    rs, _ = _lib._i32_array(res)
    P = _lib._ptr
    plans = [_lib.Plan(sets[0]["coords"]), _lib.Plan(sets[1 % sets_n]["coords"])] if planned else None
    side = torch.cuda.Stream(device=dev)
    binned = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def bin_samples(i, stream):
        """Re-bin the samples of step i into plan i % 2 on `stream` (3 small kernels, allocation reused)."""
        st = ctypes.c_void_p(stream.cuda_stream)
        _lib._check(lib.shacira_plan_rebuild(plans[i % 2].handle, 3, P(sets[i % sets_n]["coords"]), S, 0, st))

    def step(i, exchange=True):
        s = sets[i % sets_n]
        cur = torch.cuda.current_stream(dev)
        st = ctypes.c_void_p(cur.cuda_stream)
        arena.zero_()   # one memset: table gradient + decoder / density-model gradients + the bits
        if entropy and r1 > r0:
            # this rank's slice of the table on its own stream, beside the grid kernels (it only reads the table)
            ent_stream.wait_stream(cur)
            with torch.cuda.stream(ent_stream):
                bits, _, gp = _lib.entropy_bits(lat[r0:r1], None, prob, 1, None, want_grads=True, want_latent_grads=False)
                gprob.grad.copy_(gp)
                gbits.grad.copy_(bits[:1].to(torch.float32))
        if planned:
            plan = plans[i % 2]
            if overlap_binning and i > 0 and step.prebinned == i:
                cur.wait_event(binned[i % 2])           # binned on the side stream during the previous exchange
            else:
                bin_samples(i, cur)
            _lib._check(lib.shacira_latent_forward_planned_z(plan.handle, P(lat), fi, rs, L, BW, C, F, 1, P(A), P(shift), 0,
                                                             P(feats), P(z), st))
            _lib._check(lib.shacira_latent_backward_planned_z(plan.handle, P(s["g"]), P(z), fi, rs, L, BW, C, F, P(A), 0,
                                                              T, 0, P(glat.grad), P(gA.grad), P(gS.grad), st))
            if overlap_binning:
                # step i + 1's samples do not depend on this step's parameters: bin them beside the exchange. The other
                # plan was last used by step i - 1, whose kernels are ordered before this point on `cur`.
                freed[(i + 1) % 2].record(cur)
                side.wait_event(freed[(i + 1) % 2])
                bin_samples(i + 1, side)
                binned[(i + 1) % 2].record(side)
                step.prebinned = i + 1
        else:
            _lib._check(lib.shacira_latent_forward(3, P(s["coords"]), S, P(lat), fi, rs, L, BW, C, F, 1, P(A), P(shift), 0,
                                                   P(feats), P(z), st))
            _lib._check(lib.shacira_latent_backward(3, P(s["coords"]), S, P(s["g"]), P(z), fi, rs, L, BW, C, F, P(A), 0, T, 0,
                                                    P(glat.grad), P(gA.grad), P(gS.grad), st))
        if entropy and r1 > r0:
            cur.wait_stream(ent_stream)
        return arena.allreduce() if exchange else 0

    step.prebinned = -1

    def timed(exchange):
        step.prebinned = -1
        for i in range(warmup):
            step(i, exchange)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(warmup, warmup + steps):
            ncoll = step(i, exchange)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, ncoll

    ms, ncoll = timed(True)
    ms_compute = ms
    if world > 1:
        ms_compute, _ = timed(False)      # the same step without the exchange: the difference is the exposed collective
    if planned:
        for p in plans:
            p.close()
    nbytes = arena.flat.numel() * 4
    if isinstance(arena, (peer.PeerArena, peer.McArena)) and arena.timed_out():
        exchange_kind += " -- A BARRIER TIMED OUT: numbers of this run are invalid"
    if isinstance(arena, (peer.PeerArena, peer.McArena)):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        arena.close()
    bps = algorithmic_bytes_per_sample()
    return {"workload": "BASELINE cfg4 NeRF-shape ray-batch DP step (re-bin + grid fwd + bwd + grad all-reduce)",
            "n_gpus": world, "scaling": "weak", "rays_per_rank": RAYS, "samples_per_rank": S,
            "ms_per_step": ms, "samples_per_s": S * world / ms * 1e3, "rays_per_s": RAYS * world / ms * 1e3,
            "allreduce_bytes": nbytes, "exchange": exchange_kind,
            "bit_rate_rows_per_rank": (r1 - r0) if entropy else 0, "collectives_per_step": ncoll,
            "ms_per_step_without_exchange": ms_compute, "exposed_collective_us": max(0.0, (ms - ms_compute) * 1e3),
            "bytes_per_sample": bps, "alg_GBs_per_gpu": bps * S / ms / 1e6,
            "binning": ("next step's samples binned beside the exchange (side stream)" if (planned and overlap_binning)
                        else ("in line" if planned else "none")),
            "path": "plan re-binned per step + sorted lane-pair kernels + tile-staged coarse levels" if planned
                    else "unsorted lane-pair kernels + whole-level shared-memory coarse path"}


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    r = run(dev, rank, world, planned="--unplanned" not in sys.argv, overlap_binning="--no-overlap" not in sys.argv)
    if rank == 0:
        import bench
        peak, _ = bench.measured_peaks()
        r["frac_hbm_peak"] = r["alg_GBs_per_gpu"] / peak
        print(json.dumps(r), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
