"""Post-processing on the GPU, mirroring ``attrici.commands.derive_huss`` of the reference.

``calc_huss_weedon2010(hurs, ps, tas)`` has the reference's name, arguments and result
(attrici/commands/derive_huss.py:27-99): specific humidity from (counterfactual) relative humidity [%], air
pressure [Pa] and temperature [K], elementwise in float64, clipped to [1e-7, 1e-1].  One kernel of
``libattrici_b200`` (``csrc/huss.cu``); no CPU fallback.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import AttriciCudaError


def calc_huss_weedon2010(hurs_in, ps_in, tas_in, device=None):
    """Specific humidity [kg kg-1]; returns a CUDA tensor if the inputs are CUDA tensors, else a numpy array."""
    if not torch.cuda.is_available():
        raise AttriciCudaError("attrici_b200 needs a CUDA device; there is no CPU fallback")
    lib = _lib.load()
    on_device = all(isinstance(a, torch.Tensor) and a.is_cuda for a in (hurs_in, ps_in, tas_in))
    dev = hurs_in.device if on_device else torch.device(device if device is not None else "cuda")

    def prep(a):
        t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=np.float64))
        return t.to(device=dev, dtype=torch.float64)

    h, p, t = torch.broadcast_tensors(prep(hurs_in), prep(ps_in), prep(tas_in))
    h, p, t = h.contiguous(), p.contiguous(), t.contiguous()
    out = torch.empty_like(h)
    with torch.cuda.device(dev):
        _lib.check(lib.attrici_derive_huss(C.c_void_p(h.data_ptr()), C.c_void_p(p.data_ptr()), C.c_void_p(t.data_ptr()),
                                           h.numel(), C.c_void_p(out.data_ptr()),
                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out if on_device else out.cpu().numpy()
