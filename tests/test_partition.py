"""Host logic of the multi-GPU path on CPU: the reference's cell partition rule and the final gather
over a world_size-2 gloo group."""

import os

import numpy as np
import pytest

from attrici_b200.detrend import gather_cells, get_task_indices, scale_predictor
from attrici_b200.solver import cell_seeds
from oracle import attrici_oracle as oracle


def reference_rule(n, k):
    """attrici/detrend.py:138-153 restated with plain integers."""
    if n % k == 0:
        return [n // k] * k
    calls = [n // k + 1] * k
    out, tot = [], 0
    for c in calls:
        if tot + c > n:
            out.append(n - tot)
            out.extend([0] * (k - len(out)))
            break
        out.append(c)
        tot += c
    return out


@pytest.mark.parametrize("n,k", [(67420, 8), (10000, 8), (10000, 3), (7, 4), (5, 8), (16, 4), (1, 1), (100, 7)])
def test_partition_rule(n, k):
    counts = reference_rule(n, k)
    seen = []
    for r in range(k):
        idx = get_task_indices(n, r, k)
        assert len(idx) == counts[r]
        seen.extend(idx.tolist())
    assert seen == list(range(n))  # contiguous, disjoint, complete
    if (n, k) == (67420, 8):
        assert counts == [8428] * 7 + [8424]


def test_partition_errors():
    with pytest.raises(ValueError):
        get_task_indices(10, 4, 4)
    with pytest.raises(ValueError):
        get_task_indices(10, -1, 4)


def test_predictor_and_seeds_match_oracle():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ssa_gmt.npz"))["tas"]
    assert np.array_equal(scale_predictor(g, 44925), oracle.gmt_on_time_axis(g, 44925))
    lat, lon = np.array([50.75, -33.25, 0.25]), np.array([9.25, -70.75, 179.75])
    want = [oracle.cell_seed(3, a, b) for a, b in zip(lat, lon)]
    assert cell_seeds(3, lat, lon).tolist() == want


def test_predictor_on_datetime_axes_and_full_extrapolation():
    """attrici/detrend.py:687-707 on real datetime64 arithmetic: the default branch's
    ``(time - time.min()) / (time.max() - time.min())`` (also with gaps), and the ``full_extrapolation`` branch,
    which is ``scipy.interpolate.interp1d(..., fill_value="extrapolate")`` on float64 nanoseconds inside xarray."""
    from scipy.interpolate import interp1d

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ssa_gmt.npz"))["tas"]
    time = np.arange("1901-01-01", "2024-01-01", dtype="datetime64[D]").astype("datetime64[ns]")
    time = np.delete(time, slice(500, 730))                                  # an axis with a gap
    days = ((time - time[0]) / np.timedelta64(1, "D")).astype(np.int64)
    t_scaled = (time - time.min()) / (time.max() - time.min())               # the reference's expression
    want = np.interp(t_scaled, np.linspace(0, 1, len(g)), g)
    want = (want - want.min()) / (want.max() - want.min())
    assert np.array_equal(scale_predictor(g, days=days + 17), want)          # any origin
    # GMT on every 10th observation day, the last days extrapolated
    gmt_time, gd = time[::10], days[::10]
    g = g[: len(gd)]
    assert gd[-1] < days[-1] and len(g) == len(gd)
    x = (gmt_time - gmt_time.min()).astype(np.float64)                        # xarray's _floatize_x: ns as float64
    x_new = (time - gmt_time.min()).astype(np.float64)
    want = interp1d(x, g, kind="linear", fill_value="extrapolate", bounds_error=False)(x_new)
    want = (want - want.min()) / (want.max() - want.min())
    got = scale_predictor(g, days=days, gmt_days=gd, full_extrapolation=True)
    assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        scale_predictor(g, days=days, full_extrapolation=True)
    with pytest.raises(ValueError):
        scale_predictor(g)


def _worker(rank, world, port, n_cells, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        counts = [len(get_task_indices(n_cells, r, world)) for r in range(world)]
        idx = get_task_indices(n_cells, rank, world)
        local = {"params": np.stack([np.full(27, float(i)) for i in idx]) if len(idx) else np.empty((0, 27)),
                 "logp": idx.astype(np.float64) * 2.0,
                 "status": (idx % 3).astype(np.int32)}
        out = gather_cells(local, counts, dist)
        ok = (out["params"].shape == (n_cells, 27) and np.array_equal(out["params"][:, 0], np.arange(n_cells))
              and np.array_equal(out["logp"], 2.0 * np.arange(n_cells))
              and np.array_equal(out["status"], (np.arange(n_cells) % 3).astype(np.int32)))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_cells", [11, 8])
def test_gather_world_size_2_gloo(n_cells):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_cells) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_cells, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, True), (1, True)]


def _writer_worker(rank, world, port, n_cells, path, q):
    import torch.distributed as dist
    from scipy.io import netcdf_file

    from attrici_b200 import io as aio
    from attrici_b200.synthetic import tile_coordinates

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T = 37
        rng = np.random.default_rng(9)
        cfact = rng.normal(size=(T, n_cells))
        replaced = rng.integers(0, 4, (T, n_cells)).astype(np.uint8)
        lat, lon = tile_coordinates(3, 5)
        lat, lon = lat[:n_cells], lon[:n_cells]           # 15 grid points, the last ones without a cell
        counts = [len(get_task_indices(n_cells, r, world)) for r in range(world)]
        idx = get_task_indices(n_cells, rank, world)
        group = dist.new_group(backend="gloo")             # what a run under NCCL adds for host tensors
        n = aio.write_merged_output_sharded(
            path, {"cfact": cfact[:, idx], "replaced": replaced[:, idx]}, counts, lat, lon, np.arange(T, dtype=np.int32),
            dist, group=group, block_bytes=600, scalars={"logp": np.arange(n_cells, dtype=np.float64)},
            shared={"gmt_scaled": np.linspace(0, 1, T)})
        ok = True
        if rank == 0:
            nc = netcdf_file(path, "r", mmap=False)
            full = np.array(nc.variables["cfact"][:]).reshape(T, -1)
            rep = np.array(nc.variables["replaced"][:]).reshape(T, -1)
            ok = (n == T and np.array_equal(full[:, :n_cells], cfact) and np.isnan(full[:, n_cells:]).all()
                  and np.array_equal(rep[:, :n_cells], replaced.astype(np.int8)) and (rep[:, n_cells:] == 0).all()
                  and np.array_equal(np.array(nc.variables["logp"][:]).ravel()[:n_cells], np.arange(n_cells))
                  and np.array_equal(np.array(nc.variables["gmt_scaled"][:]), np.linspace(0, 1, T)))
            nc.close()
        else:
            ok = n is None
        dist.barrier()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_cells", [11, 14])
def test_sharded_merged_output_world_size_2_gloo(n_cells, tmp_path):
    """Final host gather to NetCDF (BASELINE configs[4]): uneven shards, several blocks of days."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + n_cells) % 2000
    path = str(tmp_path / "merged.nc")
    procs = [ctx.Process(target=_writer_worker, args=(r, 2, port, n_cells, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, True), (1, True)]
