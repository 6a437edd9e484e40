"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch on
the GPU box, gloo in the CPU tests). The reference is single-process/single-GPU (SURVEY
section 2: no distributed code at all); both sharding schemes come from BASELINE.json's north_star:

  * independent image INRs: unit i -> rank i mod world, no data-path collective
    (`shard_units`, results gathered once at the end with `gather_results`);
  * NeRF ray batches: data parallel over rays, parameters replicated, one exchange step per
    iteration -- SUM all-reduce of grad(latents) (+ the KB-sized decoder / density-model / MLP
    grads, flattened into one bucket) -- `allreduce_grads`.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_units(num_units, rank=None, world_size=None):
    """Indices of the independent units (images) this rank fits: round-robin."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return list(range(rank, num_units, world_size))


def split_rays(num_rays, rank=None, world_size=None):
    """[begin, end) of this rank's contiguous slice of a ray batch (sizes differ by at most 1)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, rem = divmod(num_rays, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_rows(num_rows, rank=None, world_size=None, align=4):
    """[begin, end) of this rank's contiguous slice of the latent table for table-wise work (the bit-rate loss is a sum
    over rows: each rank evaluates its slice, the all-reduce of the gradient arena adds the pieces -- SURVEY 8e).
    Slice starts are multiples of `align` rows (16-byte aligned float rows for the vectorised kernels)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    per = -(-num_rows // world_size)
    per = -(-per // align) * align
    begin = min(num_rows, rank * per)
    return begin, min(num_rows, begin + per)


def allreduce_grads(params, average=False, small_numel=1 << 16, group=None):
    """SUM all-reduce of .grad over ranks. Large gradients (the latent table) are reduced in
    place, each as its own collective launched in parameter order; everything smaller than
    `small_numel` is packed into one flat bucket so the step costs two collectives, not dozens.
    Returns the number of collectives issued."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    ws = dist.get_world_size(group)
    big, small = [], []
    for p in params:
        if p.grad is None:
            continue
        (big if p.grad.numel() >= small_numel else small).append(p.grad)
    handles, n = [], 0
    for g in big:
        handles.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True))
        n += 1
    flat = None
    if small:
        flat = torch.cat([g.reshape(-1) for g in small])
        handles.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True))
        n += 1
    for h in handles:
        h.wait()
    if flat is not None:
        off = 0
        for g in small:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    if average:
        for g in big + small:
            g.div_(ws)
    return n


class GradArena:
    """One flat float32 gradient buffer for a set of parameters: every `.grad` is a view into it, so the whole
    exchange step of the ray-batch data-parallel path is ONE all-reduce (the latent table's gradient and the KB-sized
    decoder / density / MLP gradients travel together; a separate small collective costs ~25 us of pure latency on
    NVSwitch). The kernels write their gradients straight into the views."""

    def __init__(self, params):
        self.params = [p for p in params]
        total = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(total, dtype=torch.float32, device=p0.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def allreduce(self, average=False, group=None):
        """SUM (or mean) over ranks, in place. Returns the number of collectives issued (0 or 1)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return 0
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.div_(dist.get_world_size(group))
        return 1


def gather_results(obj, group=None):
    """Every rank's python object on every rank (end-of-fit metrics of the image shards)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [obj]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out
