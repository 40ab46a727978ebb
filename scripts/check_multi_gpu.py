#!/usr/bin/env python
"""Multi-GPU path of the batched driver under torchrun (NCCL): every rank detrends the cell range the reference's
partition rule assigns to it, the per-cell results are gathered, rank 0 checks them against a single-rank run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/check_multi_gpu.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attrici_b200 import synthetic  # noqa: E402
from attrici_b200.detrend import Config, detrend_cells, get_task_indices  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
T, n_lat, n_lon = 3000, 9, 11          # 99 cells: uneven shards under the reference's rule
lat, lon = synthetic.tile_coordinates(n_lat, n_lon)
y = synthetic.tas_cells(n_lat * n_lon, T, seed=5, nan_fraction=0.05)
gmt, days = synthetic.smoothed_gmt(T), synthetic.time_axis(T)
res = detrend_cells(Config("tas", slab_cells=16), y, gmt, days, lat, lon)
idx = get_task_indices(y.shape[1], rank, world)
assert res["indices"].tolist() == idx.tolist()
assert res["params"].shape == (y.shape[1], 27), res["params"].shape       # gathered over all ranks
cf = torch.zeros((T, y.shape[1]), dtype=torch.float64, device=f"cuda:{local}")
cf[:, idx] = torch.from_numpy(np.nan_to_num(res["cfact"], nan=-1e30)).to(cf.device)
dist.all_reduce(cf)                                                        # disjoint columns: sum = merge
dist.barrier()
if rank == 0:
    dist.destroy_process_group()
    os.environ.pop("RANK", None)
    ref = detrend_cells(Config("tas", slab_cells=64, device="cuda:0"), y, gmt, days, lat, lon, gather=False)
    np.testing.assert_allclose(res["params"], ref["params"], rtol=0, atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(res["logp"], ref["logp"], rtol=1e-12, equal_nan=True)
    assert np.array_equal(res["status"], ref["status"])
    np.testing.assert_allclose(cf.cpu().numpy(), np.nan_to_num(ref["cfact"], nan=-1e30), rtol=1e-9)
    print(f"multi-GPU check ok: world {world}, shards {[len(get_task_indices(y.shape[1], r, world)) for r in range(world)]}")
else:
    dist.destroy_process_group()
