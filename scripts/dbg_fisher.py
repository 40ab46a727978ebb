import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
import helpers
from attrici_b200.solver import CellBatchSolver
from attrici_b200 import synthetic
T, C = 43464, 256
y = synthetic.tas_tile_torch(C, T, device='cuda:0', nan_fraction=0.05)
gmt, days = synthetic.smoothed_gmt(T), synthetic.time_axis(T)
s = CellBatchSolver('tas', device='cuda:0')
s.set_predictor(days, gmt, C)
ys, sc = s.scale(y)
fit = s.fit(ys)
npass = fit.n_pass.cpu().numpy()
nan_cells = torch.isnan(y).any(0).cpu().numpy()
print('passes', np.bincount(npass), 'nan cells', nan_cells.sum(), 'passes of nan cells', npass[nan_cells])
th = fit.params.cpu().numpy() * 0.98
n32, g32, f32 = s.nll_grad(ys, th, use_fp64=False, want_fisher=True)
n64, g64, f64 = s.nll_grad(ys, th, use_fp64=True, want_fisher=True)
d = (f32 - f64).abs().amax(dim=(1, 2)) / f64.abs().amax(dim=(1, 2))
print('fisher rel diff: max over valid cells', d[~torch.as_tensor(nan_cells)].max().item(), 'nan cells', d[torch.as_tensor(nan_cells)].cpu().numpy())
gd = (g32 - g64).abs().amax(1) / g64.abs().amax(1)
print('grad rel diff', gd.max().item(), 'nll rel', ((n32 - n64).abs() / n64.abs()).max().item())
c = int(np.argmax(d.cpu().numpy()))
print('worst cell', c, 'nan?', nan_cells[c])
blk = (f32[c] - f64[c]).abs().cpu().numpy()
i, j = np.unravel_index(np.argmax(blk), blk.shape)
print('worst entry', i, j, f32[c, i, j].item(), f64[c, i, j].item())
