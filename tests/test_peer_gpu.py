"""Peer-memory exchange kernel (csrc/peer_kernels.cuh) inside ONE process driving two GPUs (plain peer access, no IPC):
the sum lands bit-identically in both arenas and equals the float sum of the two inputs. Needs two visible GPUs with
peer access (skipped on the single-GPU test box; benchmarks/peer_check.py is the multi-process / NCCL comparison)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("numel", [4, 1000, 6098925 + 200])
def test_two_gpu_allreduce_in_one_process(lib, numel):
    if torch.cuda.device_count() < 2 or not torch.cuda.can_device_access_peer(0, 1):
        pytest.skip("needs two GPUs with peer access")
    from shacira_b200 import peer
    bufs = [peer.PeerBuffer(numel, "cuda:%d" % d) for d in range(2)]
    for r, b in enumerate(bufs):
        b.connect_local(bufs, r)
    torch.manual_seed(numel)
    for it in range(3):
        xs = [torch.randn(bufs[0].numel) for _ in range(2)]
        for b, x in zip(bufs, xs):
            b.flat.copy_(x.to(b.device))
        for d in range(2):
            torch.cuda.synchronize(d)
        for r, b in enumerate(bufs):         # both launches are asynchronous: the kernels meet at their barrier
            with torch.cuda.device(b.device):
                peer._lib._check(b.lib.shacira_peer_allreduce(b.ptr_array(), b.flags_offset, r, 2, b.numel,
                                                              peer._lib._stream()))
        for d in range(2):
            torch.cuda.synchronize(d)
        want = xs[0] + xs[1]
        assert torch.equal(bufs[0].flat.cpu(), want)
        assert torch.equal(bufs[1].flat.cpu(), want)
    for b in bufs:
        b.close()


@pytest.mark.parametrize("numel", [1000, 300000])
def test_two_ranks_on_one_gpu_meet_at_the_barriers(lib, numel):
    """The exchange kernel's protocol on ONE device: two arenas on cuda:0, the two ranks' kernels launched on two streams
    (they run concurrently: 2 CTAs per SM each; the flag spins are bounded, so a box that serialised them would fail this
    test after a few seconds, not hang). Barriers, epoch / ticket reuse over several calls, slice ownership and the
    rank-ordered sum are the same code paths as across GPUs; only the NVLink hop is missing."""
    from shacira_b200 import peer
    bufs = [peer.PeerBuffer(numel, "cuda:0") for _ in range(2)]
    for r, b in enumerate(bufs):
        b.connect_local(bufs, r)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.manual_seed(numel)
    for it in range(4):
        xs = [torch.randn(bufs[0].numel) for _ in range(2)]
        for b, x in zip(bufs, xs):
            b.flat.copy_(x.cuda())
        torch.cuda.synchronize()
        for r, b in enumerate(bufs):
            with torch.cuda.stream(streams[r]):
                peer._lib._check(b.lib.shacira_peer_allreduce(b.ptr_array(), b.flags_offset, r, 2, b.numel,
                                                              peer._lib._stream()))
        torch.cuda.synchronize()
        assert not bufs[0].timed_out() and not bufs[1].timed_out()
        want = xs[0] + xs[1]
        assert torch.equal(bufs[0].flat.cpu(), want) and torch.equal(bufs[1].flat.cpu(), want)
    for b in bufs:
        b.close()


def test_peer_allreduce_rejects_bad_arguments(lib):
    import ctypes
    from shacira_b200 import _lib, peer
    b = peer.PeerBuffer(16, "cuda:0")
    arr = (ctypes.c_void_p * 3)(b.ptr, b.ptr, b.ptr)
    assert lib.load().shacira_peer_allreduce(arr, b.flags_offset, 0, 3, 16, None) == _lib.ERR_UNSUPPORTED
    arr2 = (ctypes.c_void_p * 2)(b.ptr, b.ptr)
    assert lib.load().shacira_peer_allreduce(arr2, b.flags_offset, 0, 2, 6, None) == _lib.ERR_INVALID_ARGUMENT
    assert lib.load().shacira_peer_allreduce(arr2, b.flags_offset, 2, 2, 16, None) == _lib.ERR_INVALID_ARGUMENT
    b.close()
