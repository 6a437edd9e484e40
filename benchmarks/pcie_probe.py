"""PCIe probe: H2D, D2H and concurrent bidirectional copy rates with pinned memory (context for bench.py's e2e)."""
import torch
n = 64 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e2 = torch.cuda.Event(enable_timing=True)
    e2.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e2) / reps


def h2d():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def both():
    h2d()
    d2h()


print("H2D GB/s", n / t(h2d) / 1e6, "D2H GB/s", n / t(d2h) / 1e6, "both (sum) GB/s", 2 * n / t(both) / 1e6)
