"""Aggregate PCIe ceiling of the box: every rank (one per GPU, torchrun) copies between pinned host memory and its GPU
at the same time - H2D alone, D2H alone, both directions - with linear copies of the sizes the host pipeline moves per
slab.  Rank 0 prints one JSON line with the per-rank and the summed GB/s.  Single process: python scripts/micro/pcie_ranks.py

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/micro/pcie_ranks.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from attrici_b200.detrend import bind_to_gpu_cpus  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", 1))
rank = int(os.environ.get("RANK", 0))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
bound = bind_to_gpu_cpus(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
T, S = 43464, 2048
dev = f"cuda:{local}"
h_y = torch.empty(T * S, dtype=torch.float32, pin_memory=True)          # one slab of input (356 MB)
h_c = torch.empty(T * S, dtype=torch.float64, pin_memory=True)          # one slab of float64 output (712 MB)
d_y = torch.empty(T * S, dtype=torch.float32, device=dev)
d_c = torch.empty(T * S, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, n=6):
    fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    sync()
    return (time.perf_counter() - t0) / n


def h2d():
    with torch.cuda.stream(s1):
        d_y.copy_(h_y, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_c.copy_(d_c, non_blocking=True)


out = {}
for name, fn, nbytes in (("h2d", h2d, T * S * 4), ("d2h", d2h, T * S * 8), ("both", lambda: (h2d(), d2h()), T * S * 12)):
    t = timed(fn)
    rate = torch.tensor([nbytes / t / 1e9], dtype=torch.float64, device=dev)
    if world > 1:
        rates = [torch.zeros_like(rate) for _ in range(world)]
        dist.all_gather(rates, rate)
        rates = [float(r.item()) for r in rates]
    else:
        rates = [float(rate.item())]
    out[name] = {"per_rank_gbs": [round(r, 2) for r in rates], "sum_gbs": round(sum(rates), 2)}
if rank == 0:
    print(json.dumps({"n_gpus": world, "cpus_bound_per_rank": bound, "host_cpus": os.cpu_count(),
                      "copy": "linear, pinned host memory, 356 MB up / 712 MB down per copy, all ranks at once", **out}), flush=True)
if world > 1:
    dist.destroy_process_group()
