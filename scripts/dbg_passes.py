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
nan_cells = torch.isnan(y).any(0).cpu().numpy()
print('MMA' if os.environ.get('ATTRICI_NO_MMA') != '1' else 'FMA', 'passes hist', np.bincount(npass))
slow = np.flatnonzero(npass > 6)
print('slow cells', slow[:20], 'nan?', nan_cells[slow][:20], 'passes', npass[slow][:20])
print('nan cells passes', npass[nan_cells])
np.save('gpurun_out/params_%s.npy' % ('mma' if os.environ.get('ATTRICI_NO_MMA') != '1' else 'fma'), fit.params.cpu().numpy())
np.save('gpurun_out/logp_%s.npy' % ('mma' if os.environ.get('ATTRICI_NO_MMA') != '1' else 'fma'), fit.logp.cpu().numpy())
