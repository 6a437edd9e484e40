"""Peer-memory exchange step (shacira_b200/peer.py, csrc/peer_kernels.cuh) against NCCL on the NeRF arena (6.1 M-row table
gradient + decoder / density-model gradients): values, cross-rank bit-identity, the Adam-fused pass against
torch.optim.Adam, and device time of both forms.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/peer_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shacira_b200 import dp, peer  # noqa: E402


def timed(fn, dev, world, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) * 1e3


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    T = int(os.environ.get("T", "6098925"))
    mk = lambda *s: torch.nn.Parameter(torch.zeros(s, device=dev))
    params = [mk(T, 1), mk(16, 1, 4), mk(16, 4), mk(4, 3, 1), mk(1)]
    arena = peer.PeerArena(params)
    ref = dp.GradArena([mk(T, 1), mk(16, 1, 4), mk(16, 4), mk(4, 3, 1), mk(1)])
    out = {"world": world, "arena_bytes": arena.buf.bytes}
    torch.manual_seed(7 + rank)
    worst = 0.0
    for it in range(3):
        vals = [torch.randn_like(p) * (10.0 ** (it - 1)) for p in params]
        for p, q, v in zip(params, ref.params, vals):
            p.grad.copy_(v)
            q.grad.copy_(v)
        arena.allreduce()
        ref.allreduce()
        torch.cuda.synchronize()
        scale = max(float(q.grad.abs().max()) for q in ref.params)
        for p, q in zip(params, ref.params):
            worst = max(worst, float((p.grad - q.grad).abs().max()) / scale)
        # bit-identical on every rank
        chk = arena.flat.view(torch.int32).to(torch.int64).sum().reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        assert all(int(c) == int(allc[0]) for c in allc), "ranks disagree"
    out["max_rel_diff_vs_nccl"] = worst
    assert worst <= 1e-6, worst
    out["peer_allreduce_us"] = timed(arena.allreduce, dev, world)
    out["nccl_allreduce_us"] = timed(ref.allreduce, dev, world)
    # the NVSwitch multicast form (NVLS), where the fabric has it
    mparams = [mk(T, 1), mk(16, 1, 4), mk(16, 4), mk(4, 3, 1), mk(1)]
    mc = peer.McArena.try_create(mparams)
    out["multicast"] = mc is not None
    if mc is not None:
        torch.manual_seed(70 + rank)
        for it in range(2):
            vals = [torch.randn_like(p) for p in mparams]
            for p, q, v in zip(mparams, ref.params, vals):
                p.grad.copy_(v)
                q.grad.copy_(v)
            mc.allreduce()
            ref.allreduce()
            torch.cuda.synchronize()
            # the switch adds in its own order: compare against the arena's magnitude, not a one-element tensor's own sum
            scale = max(float(q.grad.abs().max()) for q in ref.params)
            w = max(float((p.grad - q.grad).abs().max()) for p, q in zip(mparams, ref.params)) / scale
            assert w <= 1e-6, w
            chk = mc.flat.view(torch.int32).to(torch.int64).sum().reshape(1)
            allc = [torch.zeros_like(chk) for _ in range(world)]
            dist.all_gather(allc, chk)
            assert all(int(c) == int(allc[0]) for c in allc), "ranks disagree (multicast)"
        out["multimem_allreduce_us"] = timed(mc.allreduce, dev, world)
        torch.cuda.synchronize()
        dist.barrier()
        mc.close()
    # graph capture of the peer exchange
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        arena.allreduce()
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10):
                arena.allreduce()
    dist.barrier()
    out["peer_allreduce_graph_us"] = timed(g.replay, dev, world, iters=10, warm=2) / 10

    # Adam inside the exchange: 3 steps against torch.optim.Adam on the NCCL-reduced gradient
    torch.manual_seed(99)
    w0 = torch.randn(T, 1, device=dev)
    table = peer.PeerTable(w0, arena.buf.numel)
    wref = torch.nn.Parameter(w0.clone())
    opt = torch.optim.Adam([wref], lr=2e-2, eps=1e-8, weight_decay=0.0)
    torch.manual_seed(1000 + rank)
    for it in range(3):
        gl = torch.randn(T, 1, device=dev)
        small = torch.randn(16, 1, 4, device=dev)
        arena.zero_()
        params[0].grad.copy_(gl)
        params[1].grad.copy_(small)
        gsum = gl.clone()
        ssum = small.clone()
        dist.all_reduce(gsum)
        dist.all_reduce(ssum)
        arena.allreduce_adam(table, lr=2e-2)
        wref.grad = gsum
        opt.step()
        torch.cuda.synchronize()
        assert float(params[0].grad.abs().max()) == 0.0                       # table gradient consumed and cleared
        assert float((params[1].grad - ssum).abs().max()) <= 1e-5 * float(ssum.abs().max())
    out["adam_max_abs_diff_vs_torch"] = float((table.data - wref.data).abs().max())
    assert out["adam_max_abs_diff_vs_torch"] <= 2e-5, out
    chk = table.data.view(torch.int32).to(torch.int64).sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    assert all(int(c) == int(allc[0]) for c in allc), "tables diverged"
    out["adam_state_floats_per_rank"] = int(table.m.numel())
    out["peer_allreduce_adam_us"] = timed(lambda: arena.allreduce_adam(table, lr=2e-2), dev, world)
    if rank == 0:
        print(json.dumps(out), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    table.close()
    arena.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
