"""PCIe copy rates with pinned host memory: contiguous vs strided (cell-slab) copies, alone and both directions at once."""
import time
import torch

T, C, S = 43464, 10000, 2048
dev = "cuda:0"
h_y = torch.empty((T, C), dtype=torch.float32, pin_memory=True)
h_c = torch.empty((T, C), dtype=torch.float64, pin_memory=True)
d_y = torch.empty((T, S), dtype=torch.float32, device=dev)
d_c = torch.empty((T, S), dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, n=3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def h2d_slab():
    with torch.cuda.stream(s1):
        d_y.copy_(h_y[:, :S], non_blocking=True)


def d2h_slab():
    with torch.cuda.stream(s2):
        h_c[:, :S].copy_(d_c, non_blocking=True)


def both():
    h2d_slab()
    d2h_slab()


hy_flat = torch.empty(T * S, dtype=torch.float32, pin_memory=True)
hc_flat = torch.empty(T * S, dtype=torch.float64, pin_memory=True)


def h2d_flat():
    with torch.cuda.stream(s1):
        d_y.view(-1).copy_(hy_flat, non_blocking=True)


def d2h_flat():
    with torch.cuda.stream(s2):
        hc_flat.copy_(d_c.view(-1), non_blocking=True)


for name, fn, nbytes in (("H2D slab (pitch)", h2d_slab, T * S * 4), ("D2H slab (pitch)", d2h_slab, T * S * 8),
                         ("H2D contiguous", h2d_flat, T * S * 4), ("D2H contiguous", d2h_flat, T * S * 8),
                         ("both slab", both, T * S * 12), ("both contiguous", lambda: (h2d_flat(), d2h_flat()), T * S * 12)):
    fn()
    t = timed(fn)
    print(f"{name:20s} {nbytes / t / 1e9:7.2f} GB/s  ({t * 1e3:.2f} ms)")
