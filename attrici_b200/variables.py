"""Variable -> model table: what ``attrici.variables.create_variable`` (attrici/variables.py:801-834)
and each class's ``__init__`` / ``scale`` / ``create_model`` / ``quantile_mapping`` / ``rescale`` decide,
flattened into the descriptor the CUDA library takes (``attrici_variable_t``).

The reference keeps these as classes (``Tas``, ``Pr``, ...: attrici/variables.py:306-798); their
arithmetic runs on the device here (``csrc/prep.cu``, ``csrc/qmap_impl.cuh``), so only the table is
left on the host.
"""

from dataclasses import dataclass

import numpy as np

from . import _lib

# distribution parameter names in the field order of attrici/distributions.py dataclasses
# (Normal :64-90, Gamma :132-158, Beta :161-187, Weibull :190-216, BernoulliGamma :93-129)
FAMILY_PARAMETERS = {
    _lib.NORMAL: ("mu", "sigma"),
    _lib.GAMMA: ("mu", "nu"),
    _lib.BETA: ("mu", "phi"),
    _lib.WEIBULL: ("alpha", "beta"),
    _lib.BERNOULLI_GAMMA: ("p", "mu", "nu"),
}

FAMILY_NAMES = {
    _lib.NORMAL: "Normal",
    _lib.GAMMA: "Gamma",
    _lib.BETA: "Beta",
    _lib.WEIBULL: "Weibull",
    _lib.BERNOULLI_GAMMA: "BernoulliGamma",
}


@dataclass(frozen=True)
class VariableSpec:
    name: str
    family: int
    scaling: int
    post_rule: int = _lib.POST_NONE
    lower_threshold: float = float("nan")  # values <= this are masked (mask_thresholded, variables.py:40-68)
    upper_threshold: float = float("nan")  # values >= this are masked
    lower_bound: float = None  # Variable.validate / check_bounds (variables.py:17-37)
    upper_bound: float = None
    scaling_keys: tuple = ()  # keys of Variable.scaling, i.e. the scaling_<k> variables of the trace file
    units: object = None  # expected ``units`` attribute of the input (check_units, variables.py:134-156): str or set

    def descriptor(self, modes):
        if not 1 <= modes <= _lib.MAX_MODES:
            raise ValueError(f"modes must be in 1..{_lib.MAX_MODES}, got {modes}")
        return _lib.Variable(self.family, self.scaling, self.post_rule, modes, self.lower_threshold,
                             self.upper_threshold)

    @property
    def parameters(self):
        return FAMILY_PARAMETERS[self.family]

    def n_params(self, modes):
        return _lib.num_params(self.family, modes)


# attrici/variables.py: Tas :306-338, Pr :341-486, Rlds :489-521, Ps :524-555, Hurs :558-609,
# Tasskew :612-657, Rsds :660-702, Tasrange :705-750, Wind :753-798
VARIABLES = {
    "tas": VariableSpec("tas", _lib.NORMAL, _lib.SCALE_UNITY, lower_bound=0.0, scaling_keys=("datamin", "scale"), units="K"),
    "ps": VariableSpec("ps", _lib.NORMAL, _lib.SCALE_UNITY, lower_bound=0.0, scaling_keys=("datamin", "scale"), units="Pa"),
    "rlds": VariableSpec("rlds", _lib.NORMAL, _lib.SCALE_UNITY, lower_bound=0.0, scaling_keys=("datamin", "scale"), units="W m-2"),
    "rsds": VariableSpec("rsds", _lib.NORMAL, _lib.SCALE_UNITY, _lib.POST_LE0_NAN, lower_bound=0.0,
                         scaling_keys=("datamin", "scale"), units="W m-2"),
    "tasskew": VariableSpec("tasskew", _lib.NORMAL, _lib.SCALE_NONE, _lib.POST_OUTSIDE01_NAN, 0.0001, 0.9999,
                            lower_bound=0.0, upper_bound=1.0, units=frozenset({"1", "K"})),
    "pr": VariableSpec("pr", _lib.BERNOULLI_GAMMA, _lib.SCALE_PR, lower_bound=0.0, scaling_keys=("scale",), units="kg m-2 s-1"),
    "tasrange": VariableSpec("tasrange", _lib.GAMMA, _lib.SCALE_RANGE, lower_threshold=0.01, lower_bound=0.0,
                             scaling_keys=("scale",), units="K"),
    "hurs": VariableSpec("hurs", _lib.BETA, _lib.SCALE_PERCENT, lower_threshold=0.01, upper_threshold=99.99,
                         lower_bound=0.0, upper_bound=100.0, scaling_keys=("scale",), units="%"),
    "sfcWind": VariableSpec("sfcWind", _lib.WEIBULL, _lib.SCALE_RANGE, lower_threshold=0.01, lower_bound=0.0,
                            scaling_keys=("scale",), units="m s-1"),
}
VARIABLES["wind"] = VariableSpec("wind", *[getattr(VARIABLES["sfcWind"], f) for f in
                                           ("family", "scaling", "post_rule", "lower_threshold", "upper_threshold",
                                            "lower_bound", "upper_bound", "scaling_keys", "units")])


def create_variable(name):
    """Counterpart of ``attrici.variables.create_variable`` (variables.py:801-834): same names, same
    ``ValueError`` for an unknown variable."""
    try:
        return VARIABLES[name]
    except KeyError:
        raise ValueError(f"Variable {name} not supported") from None


def check_units(spec, units):
    """``check_units`` of the reference (attrici/variables.py:134-156) for the ``units`` attribute of an input
    file: same ``ValueError`` texts; ``units=None`` (no attribute) passes, as there (it only logs a warning)."""
    if units is None or spec.units is None:
        return
    if isinstance(spec.units, str):
        if units != spec.units:
            raise ValueError(f"Units of data are {units}, but expected {spec.units}.")
    elif units not in spec.units:
        raise ValueError(f"Units of data are {units}, but expected one of {set(spec.units)}.")


def check_bounds(data, lower=None, upper=None, skip_inf=False):
    """``check_bounds`` of the reference (attrici/variables.py:17-37): ``ValueError(value, "is smaller than lower
    bound", lower, ".")`` / ``(value, "is bigger than upper bound", upper, ".")``.

    The reference checks one cell at a time with xarray's NaN-skipping ``min`` / ``max``; here ``data`` may be a whole
    ``[T, C]`` batch (NumPy array or CPU tensor), reduced with ``fmin`` / ``fmax`` (NaN-skipping, no copy; an all-NaN
    or empty batch passes, as ``nan < bound`` is false).  ``skip_inf``: for the input check — the reference has set
    infinite values to NaN before it builds the variable (attrici/detrend.py:311)."""
    v = data.numpy() if hasattr(data, "numpy") else np.asarray(data)
    if v.size == 0 or (lower is None and upper is None):
        return

    def extreme(ufunc):
        m = ufunc.reduce(v, axis=None)
        if skip_inf and np.isinf(m):
            finite = v[np.isfinite(v)]
            m = ufunc.reduce(finite) if finite.size else np.nan
        return m

    if lower is not None:
        m = extreme(np.fmin)
        if m < lower:
            raise ValueError(m, "is smaller than lower bound", lower, ".")
    if upper is not None:
        m = extreme(np.fmax)
        if m > upper:
            raise ValueError(m, "is bigger than upper bound", upper, ".")
