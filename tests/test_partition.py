"""Host logic of the multi-GPU path on CPU: the reference's cell partition rule and the final gather
over a world_size-2 gloo group."""

import os

import numpy as np
import pytest

from attrici_b200.detrend import (Config, detrend_cells, fit_window_indices, gather_cells, get_task_indices,
                                  scale_predictor, select_cells, stack_latlon)
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


def _reference_function(name, path="/root/reference/attrici/detrend.py"):
    """The unmodified source text of one top-level function of the reference, executed with NumPy and a silent
    logger (the module itself needs xarray / tomlkit); None where the reference is not present."""
    import ast

    if not os.path.exists(path):
        return None
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)

    class _Quiet:
        def __getattr__(self, _):
            return lambda *a, **k: None

    ns = {"np": np, "logger": _Quiet()}
    exec(compile(ast.get_source_segment(src, node), path, "exec"), ns)
    return ns[name]


@pytest.mark.skipif(not os.path.exists("/root/reference/attrici/detrend.py"), reason="reference not present")
def test_partition_rule_against_the_reference_function():
    """``get_task_indices`` against the reference's own function (attrici/detrend.py:119-169) for every task of many
    (cells, tasks) pairs, including more tasks than cells and the global land count."""
    ref = _reference_function("get_task_indices")
    rng = np.random.default_rng(0)
    pairs = [(67420, 8), (67420, 7), (1, 1), (3, 8), (8, 8), (9, 8)] + [
        (int(rng.integers(1, 3000)), int(rng.integers(1, 40))) for _ in range(150)]
    for n, k in pairs:
        for task in range(k):
            assert np.array_equal(get_task_indices(n, task, k), ref(n, task, k)), (n, k, task)
    for bad in (-1, 4):
        with pytest.raises(ValueError):
            ref(10, bad, 4)
        with pytest.raises(ValueError):
            get_task_indices(10, bad, 4)


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


def test_cell_selection_by_list_and_mask():
    """attrici/detrend.py:672-685: stacking order, ``--cells`` and the mask file."""
    glat, glon = np.array([50.75, 51.25, 51.75]), np.array([9.25, 9.75])
    lat, lon = stack_latlon(glat, glon)
    assert lat.tolist() == [50.75, 50.75, 51.25, 51.25, 51.75, 51.75] and lon.tolist() == [9.25, 9.75] * 3
    cube = np.arange(4 * 3 * 2).reshape(4, 3, 2)                                  # (time, lat, lon)
    assert np.array_equal(cube.reshape(4, -1)[:, 3], cube[:, 1, 1])               # column i * n_lon + j
    assert select_cells(lat, lon).tolist() == list(range(6))
    assert select_cells(lat, lon, cells=[(51.75, 9.25), (50.75, 9.75)]).tolist() == [4, 1]     # order as given
    with pytest.raises(KeyError):
        select_cells(lat, lon, cells=[(51.75, 9.25), (52.0, 9.75)])
    mask = np.array([[0, 1], [1, 1], [np.nan, 2]])                                # only == 1 counts
    assert select_cells(lat, lon, mask=mask, mask_lat=glat, mask_lon=glon).tolist() == [1, 2, 3]
    # a mask on a different (larger) grid: stacked order of the mask, cells outside the input are an error
    big_lat, big_lon = np.array([50.25, 50.75, 51.25]), np.array([9.25, 9.75])
    m = np.zeros((3, 2))
    m[1, 1] = m[2, 0] = 1
    assert select_cells(lat, lon, mask=m, mask_lat=big_lat, mask_lon=big_lon).tolist() == [1, 2]
    m[0, 0] = 1
    with pytest.raises(KeyError):
        select_cells(lat, lon, mask=m, mask_lat=big_lat, mask_lon=big_lon)
    # cells first, then the mask on what is left (a masked-in cell that was not selected is an error)
    assert select_cells(lat, lon, cells=[(51.25, 9.75), (50.75, 9.75)], mask=np.array([[0, 1], [0, 1], [0, 0]]),
                        mask_lat=glat, mask_lon=glon).tolist() == [1, 3]
    with pytest.raises(KeyError):
        select_cells(lat, lon, cells=[(51.25, 9.75)], mask=mask, mask_lat=glat, mask_lon=glon)
    with pytest.raises(ValueError):
        select_cells(lat, lon, mask=np.ones((2, 2)), mask_lat=glat, mask_lon=glon)


def test_fit_window_from_dates():
    """``subset_times`` of attrici/detrend.py:709-722: both ends inclusive, None = the whole series."""
    import datetime

    time = np.arange("1901-01-01", "1911-01-01", dtype="datetime64[D]").astype("datetime64[ns]")
    assert fit_window_indices(time) == (0, len(time))
    i0, i1 = fit_window_indices(time, "1902-01-01", datetime.date(1905, 12, 31))
    want = np.flatnonzero((time >= np.datetime64("1902-01-01")) & (time <= np.datetime64("1905-12-31")))
    assert (i0, i1) == (want[0], want[-1] + 1) and i1 - i0 == 4 * 365 + 1
    assert fit_window_indices(time, None, "2021-12-31") == (0, len(time))     # the reference tests' stop date
    with pytest.raises(ValueError):
        fit_window_indices(time, "1950-01-01", None)
    with pytest.raises(TypeError):
        fit_window_indices(np.arange(10))


def test_driver_rejects_what_the_reference_rejects():
    """attrici/detrend.py:646-647 (modes xor window_size), the window-size rule and the window model's need of a
    calendar axis are raised before anything touches the GPU."""
    y = np.zeros((10, 2), dtype=np.float32)
    args = (y, np.linspace(0, 1, 10), np.arange(10), [0.25, 0.25], [0.25, 0.75])
    with pytest.raises(ValueError, match="Exactly one of `modes` and `window_size` must be set"):
        detrend_cells(Config("tas", modes=4, window_size=31), *args)
    with pytest.raises(ValueError, match="Exactly one of `modes` and `window_size` must be set"):
        detrend_cells(Config("tas", modes=None), *args)
    with pytest.raises(TypeError, match="datetime64"):      # day of the year needs calendar days (util.py:126-133)
        detrend_cells(Config("tas", modes=None, window_size=31), *args)
    with pytest.raises(ValueError, match="Window size must be an odd number"):
        detrend_cells(Config("tas", modes=None, window_size=30), *args)
    bad_gmt = np.linspace(0, 1, 10)
    bad_gmt[3] = np.nan
    with pytest.raises(ValueError, match="GMT data must not contain NaN or infinite values"):
        detrend_cells(Config("tas"), y, bad_gmt, *args[2:])
    with pytest.raises(ValueError, match="GMT data must have dimension 'time'"):
        detrend_cells(Config("tas"), y, np.zeros((10, 2)), *args[2:])
    with pytest.raises(ValueError, match="only provides solver='cuda'"):
        detrend_cells(Config("tas", solver="scipy"), *args)
    with pytest.raises(ValueError, match="Variable huss not supported"):
        detrend_cells(Config("huss"), *args)


class _StubSolver:
    """Stands in for CellBatchSolver on the CPU: records the calls and returns arrays that encode which cells of
    which slab they came from, so that the slab bookkeeping of ``detrend._slabs_on_device`` can be checked."""

    def __init__(self):
        import torch

        self.device, self.n_params, self.calls, self.torch = torch.device("cpu"), 3, [], torch

    def set_predictor(self, days, gmt, n_cells):
        self.calls.append(("predictor", n_cells))

    def scale(self, y, fit_window=None):
        self.calls.append(("scale", tuple(y.shape), fit_window))
        t = self.torch
        return y * 0.5, t.stack([y[0].double(), t.full_like(y[0], 2.0).double()], dim=1)

    def fit(self, ys):
        t = self.torch
        cells = ys[0] * 2.0                                     # the stub's y holds the cell number in every row
        from attrici_b200.solver import FitResult

        return FitResult(t.stack([cells, cells + 0.25, cells + 0.5], dim=1).double(), -cells.double(),
                         t.ones(ys.shape[1], dtype=t.int32), t.full((ys.shape[1],), 5, dtype=t.int32))

    def quantile_map(self, y, ys, params, scaling, seeds=None, want_replaced=True, want_scaled=False):
        self.calls.append(("map", None if seeds is None else list(seeds)))
        cfact = y.double() + params[:, 0][None, :] * 1000.0 + scaling[:, 1][None, :]
        return cfact, (y > 3).to(self.torch.uint8), (cfact - 1.0 if want_scaled else None)

    def eval_parameter(self, params, which, counterfactual=False):
        base = {"mu": 1.0, "sigma": 2.0}[which] + (0.5 if counterfactual else 0.0)
        return (params[:, 0] + base)[None, :].repeat(6, 1)


def test_slab_loop_of_the_fit_only_and_trace_modes():
    import torch

    from attrici_b200 import detrend as D

    T, n = 6, 7
    y = torch.arange(n, dtype=torch.float32)[None, :].repeat(T, 1)
    seeds = np.arange(100, 100 + n, dtype=np.uint32)
    # fit only: three slabs (3 + 3 + 1), no map
    s = _StubSolver()
    out = D._slabs_on_device(s, D.Config("tas", slab_cells=3), y, np.arange(T), np.zeros(T), (1, 5), seeds, None, False)
    assert out["cfact"] is None and out["replaced"] is None
    assert out["params"][:, 0].tolist() == list(range(n)) and out["logp"].tolist() == [-float(i) for i in range(n)]
    assert out["status"].tolist() == [1] * n and out["n_pass"].tolist() == [5] * n
    assert out["scaling"][:, 0].tolist() == list(range(n))
    assert [c for c in s.calls if c[0] == "scale"] == [("scale", (T, 3), (1, 5)), ("scale", (T, 3), (1, 5)), ("scale", (T, 1), (1, 5))]
    assert not [c for c in s.calls if c[0] == "map"]
    # given trace: no fit, the trace's coefficients and scaling reach the map of the right cells, seeds sliced alike
    s = _StubSolver()
    trace = {"params": np.arange(n)[:, None] + np.zeros((n, 3)) + 10.0, "logp": np.arange(n) * 1.5,
             "scaling": np.column_stack([np.zeros(n), np.arange(n) + 0.125])}
    out = D._slabs_on_device(s, D.Config("tas", slab_cells=4), y, np.arange(T), np.zeros(T), None, seeds, trace, True)
    want = y.double() + (torch.arange(n) + 10.0)[None, :] * 1000.0 + (torch.arange(n) + 0.125)[None, :]
    assert torch.equal(out["cfact"], want) and torch.equal(out["replaced"], (y > 3).to(torch.uint8))
    assert np.array_equal(out["params"].numpy(), trace["params"]) and np.array_equal(out["logp"].numpy(), trace["logp"])
    assert np.array_equal(out["scaling"].numpy(), trace["scaling"]) and out["n_pass"].tolist() == [0] * n
    assert [c[1] for c in s.calls if c[0] == "map"] == [[100, 101, 102, 103], [104, 105, 106]]
    # report_variables beyond cfact / replaced: every series lands in the columns of its slab
    s = _StubSolver()
    names = D.report_series(D.create_variable("tas"), ("all",))
    assert names == ["y_scaled", "cfact_scaled", "mu", "sigma", "mu_cfact", "sigma_cfact"]
    assert D.report_series(D.create_variable("pr"), ("cfact", "p_cfact", "y", "nu")) == ["nu", "p_cfact"]
    with pytest.raises(KeyError):
        D.report_series(D.create_variable("tas"), ("cfact", "nu"))
    out = D._slabs_on_device(s, D.Config("tas", slab_cells=4), y, np.arange(T), np.zeros(T), None, None, trace, True, names)
    ser = out["series"]
    assert ser["y_scaled"].dtype == torch.float32 and torch.equal(ser["y_scaled"], y * 0.5)
    assert torch.equal(ser["cfact_scaled"], out["cfact"] - 1.0)
    cells = torch.arange(n, dtype=torch.float64) + 10.0
    for k, add in (("mu", 1.0), ("sigma", 2.0), ("mu_cfact", 1.5), ("sigma_cfact", 2.5)):
        assert torch.equal(ser[k], (cells + add)[None, :].repeat(T, 1)), k
    # a trace without scaling constants keeps the ones of the data
    s = _StubSolver()
    del trace["scaling"]
    out = D._slabs_on_device(s, D.Config("tas", slab_cells=4), y, np.arange(T), np.zeros(T), None, None, trace, True)
    assert out["scaling"][:, 1].tolist() == [2.0] * n


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


@pytest.mark.parametrize("n_cells", [11, 8, 1])
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
                  and np.array_equal(rep[:, :n_cells], aio.replaced_flags(replaced), equal_nan=True)   # float flags
                  and (rep[:, n_cells:] == 0).all()
                  and np.array_equal(np.array(nc.variables["logp"][:]).ravel()[:n_cells], np.arange(n_cells))
                  and np.array_equal(np.array(nc.variables["gmt_scaled"][:]), np.linspace(0, 1, T)))
            nc.close()
        else:
            ok = n is None
        dist.barrier()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_cells", [11, 14, 1])
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


class _StubBatchSolver(_StubSolver):
    """``CellBatchSolver`` stand-in for the driver under gloo: ``detrend_host`` returns results that encode the cell
    (the first row of ``y`` holds the global cell number), the slab calls are those of ``_StubSolver``."""

    def __init__(self, spec, modes=4, device="cpu", use_fp64=False):
        super().__init__()
        self.spec, self.n_params = spec, spec.n_params(modes)

    def fit(self, ys):
        import torch

        from attrici_b200.solver import FitResult

        cells = ys[0].double() * 2.0
        params = cells[:, None] + torch.arange(self.n_params, dtype=torch.float64)[None, :] / 100.0
        n = ys.shape[1]
        return FitResult(params, -cells, torch.zeros(n, dtype=torch.int32), torch.full((n,), 5, dtype=torch.int32))

    def eval_parameter(self, params, which, counterfactual=False):
        base = {"mu": 1.0, "sigma": 2.0}[which] + (0.5 if counterfactual else 0.0)
        return (params[:, 0] + base)[None, :].repeat(self._T, 1)

    def scale(self, y, fit_window=None):
        self._T = y.shape[0]
        return super().scale(y, fit_window)

    def detrend_host(self, y, days, gmt, fit_window=None, seeds=None, slab_cells=2048, out=None, want_replaced=True):
        import torch

        cells = y[0].double()
        n = y.shape[1]
        return {"cfact": y.double() + 1000.0 * cells[None, :], "replaced": (y > 3).to(torch.uint8),
                "params": cells[:, None] + torch.arange(self.n_params, dtype=torch.float64)[None, :] / 100.0,
                "logp": -cells, "scaling": torch.stack([cells, cells + 0.5], dim=1),
                "status": torch.zeros(n, dtype=torch.int32), "n_pass": torch.full((n,), 5, dtype=torch.int32)}


def _driver_worker(rank, world, port, n_cells, q):
    import torch
    import torch.distributed as dist

    import attrici_b200.solver as solver_module

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        solver_module.CellBatchSolver = _StubBatchSolver      # the GPU stays out of this test: host logic only
        T = 6
        y = np.arange(n_cells, dtype=np.float32)[None, :].repeat(T, 0)
        lat, lon = np.arange(n_cells) * 0.5, np.zeros(n_cells)
        args = (y, np.linspace(0, 1, T), np.arange(T), lat, lon)
        idx = get_task_indices(n_cells, rank, world)
        cells = np.arange(n_cells, dtype=np.float64)
        res = detrend_cells(Config("tas", device="cpu"), *args)
        ok = [res["indices"].tolist() == idx.tolist(),
              res["cfact"].shape == (T, len(idx)) and np.array_equal(res["cfact"][0], 1001.0 * idx),
              res["params"].shape == (n_cells, 27) and np.array_equal(res["params"][:, 0], cells),   # gathered
              np.array_equal(res["logp"], -cells), np.array_equal(res["scaling"][:, 1], cells + 0.5),
              res["status"].dtype == np.int32 and res["n_pass"].tolist() == [5] * n_cells]
        local = detrend_cells(Config("tas", device="cpu"), *args, gather=False)
        ok.append(local["params"].shape == (len(idx), 27))
        # trace-file and report_variables modes: every rank slices the full-size trace by its own cell range
        trace = {"params": cells[:, None] * 3.0 + np.zeros((n_cells, 27)), "logp": cells * 7.0}
        tr = detrend_cells(Config("tas", device="cpu", slab_cells=2, report_variables=("cfact", "mu_cfact")), *args,
                           trace=trace)
        ok += [np.array_equal(tr["params"], trace["params"]), np.array_equal(tr["logp"], trace["logp"]),
               list(tr["series"]) == ["mu_cfact"] and tr["series"]["mu_cfact"].shape == (T, len(idx)),
               len(idx) == 0 or np.array_equal(tr["series"]["mu_cfact"][0], 3.0 * idx + 1.5),
               (tr["n_pass"] == 0).all()]
        fo = detrend_cells(Config("tas", device="cpu", slab_cells=3, fit_only=True), *args)
        ok += [fo["cfact"] is None, np.array_equal(fo["params"][:, 0], cells)]
        q.put((rank, [bool(x) for x in ok]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_cells", [11, 1])
def test_batched_driver_world_size_2_gloo(n_cells):
    """``detrend_cells`` under torch.distributed (gloo, two ranks) with the solver stubbed out: the reference's
    partition of the cells, local series, per-cell results gathered over uneven (or empty) shards, and the trace /
    report-variables / fit-only modes."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() + n_cells) % 2000
    procs = [ctx.Process(target=_driver_worker, args=(r, 2, port, n_cells, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(2))
    assert all(got[0]) and all(got[1]), got
