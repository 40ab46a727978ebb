"""attrici_b200: B200-native batched solver for ATTRICI's per-gridcell detrending path.

The hot path (scale -> MAP fit -> factual/counterfactual parameters -> quantile map) runs as
hand-written sm_100a CUDA kernels in ``libattrici_b200.so`` behind a C ABI (``include/attrici_b200.h``);
this package is the thin Python host layer mirroring the reference's solver plugin interface.
There is no CPU fallback: solver entry points raise if the library or a GPU is missing.
"""

__version__ = "0.1.0"

from . import _lib  # noqa: F401
from .variables import VARIABLES, create_variable  # noqa: F401


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import attrici_b200` stays cheap
    if name in ("CellBatchSolver", "cell_seeds", "FitResult"):
        from . import solver

        return getattr(solver, name)
    if name == "ModelCuda":
        from .estimation import ModelCuda

        return ModelCuda
    if name in ("detrend_cells", "get_task_indices", "Config"):
        from . import detrend

        return getattr(detrend, name)
    raise AttributeError(name)
