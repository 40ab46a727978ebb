#!/usr/bin/env python
"""Streaming bandwidth of one B200 for the read : write mixes of the streaming kernels, with library kernels as the
yardstick: copy (1 : 1, what MEASURED_PEAKS.json states), float32 -> float64 conversion (4 B in, 8 B out - the mix of
the quantile map without its flag plane), fill (write only), sum (read only).  CUDA events, 20 repetitions after 3
warm-ups, buffers far larger than L2.  Usage: python scripts/micro/mix_bw.py [elements]"""
import json
import sys

import torch


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 434_640_000   # the cell-days of one 10 000-cell tile
    dev = torch.device("cuda:0")
    a32 = torch.rand(n, device=dev, dtype=torch.float32)
    b32 = torch.empty_like(a32)
    c64 = torch.empty(n, device=dev, dtype=torch.float64)
    res = {"elements": n}
    t = timed(lambda: b32.copy_(a32))
    res["copy_f32 (4 B in, 4 B out)"] = {"ms": t * 1e3, "GB/s": 8.0 * n / t / 1e9}
    t = timed(lambda: c64.copy_(a32))
    res["convert_f32_to_f64 (4 B in, 8 B out)"] = {"ms": t * 1e3, "GB/s": 12.0 * n / t / 1e9}
    t = timed(lambda: c64.fill_(1.0))
    res["fill_f64 (8 B out)"] = {"ms": t * 1e3, "GB/s": 8.0 * n / t / 1e9}
    t = timed(lambda: a32.sum())
    res["sum_f32 (4 B in)"] = {"ms": t * 1e3, "GB/s": 4.0 * n / t / 1e9}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
