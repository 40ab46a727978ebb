"""ctypes binding of ``libattrici_b200.so`` (the C ABI declared in ``include/attrici_b200.h``).

This is the stub a maintainer of the reference would add: plain pointers and
sizes, no torch types in the signatures.  There is NO CPU fallback: if the library has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C attrici_b200/csrc``) importing
this module raises, and so does every solver entry point.
"""

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libattrici_b200.so")

ABI_VERSION = 2
MAX_MODES = 4

# enums of include/attrici_b200.h
NORMAL, GAMMA, BETA, WEIBULL, BERNOULLI_GAMMA = range(5)
SCALE_UNITY, SCALE_NONE, SCALE_PERCENT, SCALE_RANGE, SCALE_PR = range(5)
POST_NONE, POST_LE0_NAN, POST_OUTSIDE01_NAN = range(3)
STATUS_CONVERGED, STATUS_RUNNING, STATUS_STALLED, STATUS_NONFINITE, STATUS_NO_DATA = range(5)
REPLACED_NONE, REPLACED_NAN, REPLACED_PINF, REPLACED_NINF = range(4)


class Variable(C.Structure):
    """``attrici_variable_t``"""

    _fields_ = [
        ("family", C.c_int32),
        ("scaling", C.c_int32),
        ("post_rule", C.c_int32),
        ("modes", C.c_int32),
        ("lower_threshold", C.c_float),
        ("upper_threshold", C.c_float),
    ]


class FitOptions(C.Structure):
    """``attrici_fit_options_t``"""

    _fields_ = [
        ("max_pass", C.c_int32),
        ("use_fp64", C.c_int32),
        ("tol_pred", C.c_double),
        ("lambda0", C.c_double),
        ("zero_init", C.c_int32),
        ("profile", C.c_int32),
    ]


class HostOptions(C.Structure):
    """``attrici_host_options_t``"""

    _fields_ = [
        ("prep_flags", C.c_int32),
        ("cfact_dtype", C.c_int32),
        ("replaced_mode", C.c_int32),
        ("slab_major", C.c_int32),
        ("sparse_capacity", C.c_int64),
    ]


class AttriciCudaError(RuntimeError):
    pass


_vp, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
_SIGNATURES = {
    "attrici_last_error": (C.c_char_p, []),
    "attrici_abi_version": (_i, []),
    "attrici_launch_count": (C.c_longlong, []),
    "attrici_fit_profile": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_double)]),
    "attrici_num_params": (_i, [_i, _i]),
    "attrici_default_fit_options": (_i, [C.POINTER(FitOptions)]),
    "attrici_workspace_bytes": (_i, [_i, _i, _i, C.POINTER(_sz)]),
    "attrici_build_basis": (_i, [_vp, _sz, _i, _vp, _vp, _i, _i, _vp]),
    "attrici_prep_scale_mask": (_i, [_vp, _sz, C.POINTER(Variable), _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp]),
    "attrici_prep_scale_mask_ex": (_i, [_vp, _sz, C.POINTER(Variable), _vp, _vp, _i64, _i, _i, _i, _i, _vp, _i, _vp]),
    "attrici_nll_grad": (_i, [_vp, _sz, C.POINTER(Variable), _vp, _i64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "attrici_fit_batch": (_i, [_vp, _sz, C.POINTER(Variable), _vp, _i64, _i, _i, _i, _i, C.POINTER(FitOptions),
                                _vp, _vp, _vp, _vp, _vp]),
    "attrici_quantile_map": (_i, [_vp, _sz, C.POINTER(Variable), _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _i64, _vp]),
    "attrici_eval_parameter": (_i, [_vp, _sz, C.POINTER(Variable), _i, _i, _vp, _i, _i, _vp, _i64, _vp]),
    "attrici_window_workspace_bytes": (_i, [_i, C.POINTER(_sz)]),
    "attrici_window_fit": (_i, [_vp, _sz, C.POINTER(Variable), _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp,
                                 _vp, _vp, _vp]),
    "attrici_window_quantile_map": (_i, [C.POINTER(Variable), _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                                          _vp]),
    "attrici_mt19937_uniform": (_i, [_vp, _i, _i, _vp, _i64, _vp]),
    "attrici_ssa_workspace_bytes": (_i, [_i, _i, C.POINTER(_sz)]),
    "attrici_ssa_first_component": (_i, [_vp, _i, _i, _vp, _vp, _sz, _i, C.c_double, _vp, _vp, _vp]),
    "attrici_bootstrap_block_choices": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "attrici_bootstrap_scores": (_i, [_vp, _sz, C.POINTER(Variable), _vp, _i64, _i, _i, _vp, _vp, _i64, _vp]),
    "attrici_bootstrap_resample": (_i, [_vp, _sz, C.POINTER(Variable), _vp, _i64, _i, _i, _vp, _vp, _i64, _vp, _i, _i,
                                         _i, _i, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "attrici_bootstrap_stats": (_i, [C.POINTER(Variable), _vp, _i64, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "attrici_derive_huss": (_i, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "attrici_detrend_host_scratch_bytes": (_i, [_i, _i, _i, C.POINTER(_sz)]),
    "attrici_detrend_host": (_i, [C.POINTER(Variable), C.POINTER(FitOptions), _vp, _i64, _i, _i, _i, _i, _vp, _vp,
                                   _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i]),
    "attrici_default_host_options": (_i, [C.POINTER(HostOptions)]),
    "attrici_host_slab_range": (_i, [_i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "attrici_detrend_host_ex": (_i, [C.POINTER(Variable), C.POINTER(FitOptions), C.POINTER(HostOptions), _vp, _i64, _i, _i,
                                      _i, _i, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i]),
    "attrici_quantile_map_f32": (_i, [_vp, _sz, C.POINTER(Variable), _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp]),
}
F64, F32 = 0, 1
FLAGS_DENSE, FLAGS_SPARSE, FLAGS_OFF = 0, 1, 2
PREP_EXACT_SUMS = 1      # attrici_prep_scale_mask_ex: two-pass sums of the float32 y_scaled
WHICH_EXPECTATION = -1   # attrici_eval_parameter: Distribution.expectation() instead of a parameter
EXPORTS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library (once) and check its ABI version; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AttriciCudaError(
            f"{LIB_PATH} not found: build it with `make -C attrici_b200/csrc` (nvcc, sm_100a). "
            "attrici_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    got = lib.attrici_abi_version()
    if got != ABI_VERSION:
        raise AttriciCudaError(f"libattrici_b200.so has ABI version {got}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc):
    """Turn a non-zero return code into an exception carrying ``attrici_last_error()``."""
    if rc != 0:
        msg = load().attrici_last_error().decode("utf-8", "replace")
        kind = "invalid argument" if rc < 0 else f"cudaError {rc}"
        raise AttriciCudaError(f"attrici_b200 call failed ({kind}): {msg}")


def num_params(family, modes):
    n = load().attrici_num_params(family, modes)
    if n < 0:
        check(n)
    return n


def default_fit_options():
    o = FitOptions()
    check(load().attrici_default_fit_options(C.byref(o)))
    return o


def workspace_bytes(family, T, C_):
    n = _sz()
    check(load().attrici_workspace_bytes(family, T, C_, C.byref(n)))
    return n.value


def detrend_host_scratch_bytes(family, T, slab_cells):
    n = _sz()
    check(load().attrici_detrend_host_scratch_bytes(family, T, slab_cells, C.byref(n)))
    return n.value


def launch_count():
    return int(load().attrici_launch_count())


def fit_profile(reset=True):
    """(summed pass-kernel ms, launches, cell-days covered) of the profiled fits of this thread."""
    ms, n, cd = C.c_double(), C.c_longlong(), C.c_double()
    check(load().attrici_fit_profile(int(reset), C.byref(ms), C.byref(n), C.byref(cd)))
    return ms.value, n.value, cd.value
