"""N>1 host logic on CPU: two processes, gloo backend, 127.0.0.1 rendezvous."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from shacira_b200 import dp
    torch.manual_seed(0)
    table = torch.nn.Parameter(torch.zeros(1 << 17, 1))        # "latents": reduced in place
    small = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(7))]
    frozen = torch.nn.Parameter(torch.zeros(5))                 # no grad: skipped
    table.grad = torch.full_like(table, float(rank + 1))
    for i, p in enumerate(small):
        p.grad = torch.full_like(p, float(10 * (i + 1) + rank))
    n = dp.allreduce_grads([table, frozen] + small)
    ok = n == 2 and bool((table.grad == 3.0).all())             # 1 + 2
    ok &= bool((small[0].grad == 21.0).all()) and bool((small[1].grad == 41.0).all())
    b, e = dp.split_rays(4096, rank, world)
    ok &= (e - b) == 2048
    # table rows for table-wise work (bit-rate loss): aligned, disjoint, covering slices; summing the slices' partial
    # results over the ranks gives the whole-table value
    r0, r1 = dp.shard_rows(6098925)
    ok &= r0 % 4 == 0 and (r1 == 6098925 or r1 % 4 == 0)
    spans = dp.gather_results((r0, r1))
    ok &= spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == 6098925
    w = torch.arange(1000, dtype=torch.float64)
    q0, q1 = dp.shard_rows(1000)
    part = torch.nn.Parameter(torch.zeros(1, dtype=torch.float64))
    part.grad = w[q0:q1].sum().reshape(1)
    dp.allreduce_grads([part], small_numel=0)
    ok &= float(part.grad) == float(w.sum())
    res = dp.gather_results({"rank": rank, "units": dp.shard_units(5)})
    ok &= [r["units"] for r in res] == [[0, 2, 4], [1, 3]]
    table.grad = torch.full_like(table, float(rank + 1))
    dp.allreduce_grads([table], average=True)
    ok &= bool((table.grad == 1.5).all())
    # GradArena: every gradient a view of one flat buffer, the exchange step is ONE collective
    a, b2 = torch.nn.Parameter(torch.zeros(1000, 2)), torch.nn.Parameter(torch.zeros(3, 5))
    arena = dp.GradArena([a, b2])
    a.grad.fill_(float(rank + 1))
    b2.grad.fill_(float(10 + rank))
    ok &= arena.allreduce() == 1 and bool((a.grad == 3.0).all()) and bool((b2.grad == 21.0).all())
    ok &= a.grad.data_ptr() == arena.flat.data_ptr() and arena.flat.numel() == 2015
    arena.zero_()
    ok &= not bool(a.grad.any()) and not bool(b2.grad.any())
    # the peer-memory arenas need CUDA IPC: on a host without a GPU every rank must AGREE to fall back (None on all
    # ranks, the gradients still attached to nothing), never one rank with an arena and one without
    from shacira_b200 import peer
    c1, c2 = torch.nn.Parameter(torch.zeros(64, 1)), torch.nn.Parameter(torch.zeros(3))
    ok &= peer.PeerArena.try_create([c1, c2]) is None and peer.McArena.try_create([c1, c2]) is None
    fallback = dp.GradArena([c1, c2])
    c1.grad.fill_(float(rank + 1))
    ok &= fallback.allreduce() == 1 and bool((c1.grad == 3.0).all())
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_allreduce_grads_world2_gloo():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
