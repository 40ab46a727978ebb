"""eval_parameter (report variables): phase-ordered against day-ordered kernel, 10 000 cells x 43 464 days."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from attrici_b200 import synthetic
from attrici_b200.solver import CellBatchSolver
T, C = synthetic.DAYS_1901_2019, 10000
s = CellBatchSolver("tas", device="cuda:0")
s.set_predictor(synthetic.time_axis(T), synthetic.smoothed_gmt(T), C)
params = torch.randn((C, s.n_params), dtype=torch.float64, device="cuda") * 0.1
out = torch.empty((T, C), dtype=torch.float64, device="cuda")
for which in ("mu", "sigma", "expectation"):
    s.eval_parameter(params, which, out=out); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): s.eval_parameter(params, which, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(os.environ.get("ATTRICI_NO_FOLD", "0"), which, f"{ms:.3f} ms  {T*C*8/ms/1e6:.0f} GB/s")
