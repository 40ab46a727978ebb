import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
from attrici_b200.solver import CellBatchSolver
from attrici_b200 import synthetic
T, C = 43464, 10000
y = synthetic.tas_tile_torch(C, T, seed=1234, device='cuda:0', nan_fraction=0.001)
gmt, days = synthetic.smoothed_gmt(T), synthetic.time_axis(T)
s = CellBatchSolver('tas', device='cuda:0')
s.set_predictor(days, gmt, C)
ys, sc = s.scale(y)
fit = s.fit(ys)
npass = fit.n_pass.cpu().numpy()
print('slow', np.flatnonzero(npass > 6), npass[npass > 6])
m = torch.nanmean(ys.double(), 0); v = torch.nanmean(ys.double() ** 2, 0) - m * m
for name, th in (('start', None), ('opt', fit.params.clone())):
    if th is None:
        th = torch.zeros((C, 27), dtype=torch.float64, device='cuda:0'); th[:, 0] = m; th[:, 18] = 0.5 * torch.log(v)
    n32, g32, f32 = s.nll_grad(ys, th, use_fp64=False, want_fisher=True)
    n64, g64, f64 = s.nll_grad(ys, th, use_fp64=True, want_fisher=True)
    gerr = (g32 - g64).abs().amax(1); ferr = (f32 - f64).abs().amax(dim=(1, 2)); nerr = (n32 - n64).abs()
    print(name, 'grad abs err: median %.3g max %.3g at %d' % (gerr.median().item(), gerr.max().item(), gerr.argmax().item()),
          '| fisher abs err median %.3g max %.3g at %d' % (ferr.median().item(), ferr.max().item(), ferr.argmax().item()),
          '| nll abs err median %.3g max %.3g at %d' % (nerr.median().item(), nerr.max().item(), nerr.argmax().item()))
    for c in (9135, 9136, 9138, 9139, 9140, 9264, 9754):
        print('   cell', c, 'passes', npass[c], 'gerr %.3g ferr %.3g nerr %.3g' % (gerr[c].item(), ferr[c].item(), nerr[c].item()), 'sigma0 %.4f' % np.exp(th[c, 18].item()))
