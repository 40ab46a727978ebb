import numpy as np, torch, sys, os
sys.path.insert(0, '.')
from attrici_b200 import synthetic
from attrici_b200.solver import CellBatchSolver
T, C = 3200, 40
nanf = float(os.environ.get("NANF", "0.1"))
y = synthetic.tas_cells(C, T, seed=3, nan_fraction=nanf)
gmt, days = synthetic.smoothed_gmt(T), synthetic.time_axis(T)
yd = torch.from_numpy(y).cuda()
for win in (None, (100, T - 50), (0, T - 50), (100, T)):
    s = CellBatchSolver("tas", device="cuda:0")
    a = s.detrend_device(yd, days, gmt, fit_window=win)
    a2 = s.detrend_device(yd, days, gmt, fit_window=win)
    pa, pb = a["fit"].params.cpu().numpy(), a2["fit"].params.cpu().numpy()
    la, lb = a["fit"].logp.cpu().numpy(), a2["fit"].logp.cpu().numpy()
    print(win, "params maxdiff", np.nanmax(np.abs(pa - pb)), "logp maxdiff", np.nanmax(np.abs(la - lb)), "npass", a["fit"].n_pass.cpu().numpy()[:6])
    th = pa.copy()
    ys, _ = s.scale(yd, win)
    n1, g1, _ = s.nll_grad(ys, th, fit_window=win)
    n2, g2, _ = s.nll_grad(ys, th, fit_window=win)
    print("   nll_grad repeat: nll diff", (n1 - n2).abs().max().item(), "grad diff", (g1 - g2).abs().max().item())
