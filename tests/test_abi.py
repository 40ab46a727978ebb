"""The C-ABI library loads and exports every symbol include/attrici_b200.h declares; argument errors are
reported as ATTRICI_E* codes with a message (no compute call here: no GPU needed)."""

import ctypes as C
import os
import re

import pytest

from attrici_b200 import _lib
from attrici_b200.variables import VARIABLES, create_variable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "attrici_b200.h")).read()
    return sorted(set(re.findall(r"ATTRICI_API\s+[\w\s\*]+?\b(attrici_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.EXPORTS) == syms  # the ctypes stub binds exactly the header


def header_prototypes():
    """name -> (return type, [argument types]) of every ATTRICI_API declaration, comments stripped."""
    src = open(os.path.join(ROOT, "include", "attrici_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"ATTRICI_API\s+([\w\s\*]+?)\b(attrici_\w+)\s*\(([^)]*)\)\s*;", src):
        args = [a.strip() for a in args.split(",")] if args.strip() not in ("", "void") else []
        out[name] = (" ".join(ret.split()), [re.sub(r"\s*\b\w+$", "", a).strip() if not a.endswith("*") else a for a in args])
    return out


def _ctype_of(c_type):
    t = c_type.replace("const ", "").strip()
    if t.endswith("*"):
        return "pointer"
    return {"int": C.c_int, "int32_t": C.c_int32, "int64_t": C.c_int64, "size_t": C.c_size_t, "double": C.c_double,
            "float": C.c_float, "long long": C.c_longlong}[t]


def test_ctypes_signatures_match_the_header_prototypes():
    """Every binding of attrici_b200/_lib.py has the argument count and the scalar / pointer kinds of its
    declaration in include/attrici_b200.h (a mismatch would corrupt the call silently under ctypes)."""
    protos = header_prototypes()
    assert sorted(protos) == sorted(_lib._SIGNATURES)
    for name, (res, args) in _lib._SIGNATURES.items():
        ret, c_args = protos[name]
        assert len(args) == len(c_args), (name, len(args), c_args)
        for i, (bound, declared) in enumerate(zip(args, c_args)):
            kind = _ctype_of(declared)
            if kind == "pointer":
                assert bound is C.c_void_p or bound is C.c_char_p or issubclass(bound, C._Pointer), (name, i, declared)
            else:
                assert C.sizeof(bound) == C.sizeof(kind) and not issubclass(bound, C._Pointer) and bound is not C.c_void_p, \
                    (name, i, declared, bound)
        if ret.endswith("*"):
            assert res is C.c_char_p
        else:
            assert C.sizeof(res) == C.sizeof(_ctype_of(ret)), (name, ret)


def test_abi_version_and_param_counts():
    lib = _lib.load()
    assert lib.attrici_abi_version() == _lib.ABI_VERSION
    # P = 3 + 6m (27 at m = 4), Bernoulli-Gamma 5 + 10m (45)
    for m in (1, 2, 3, 4):
        for fam in (_lib.NORMAL, _lib.GAMMA, _lib.BETA, _lib.WEIBULL):
            assert _lib.num_params(fam, m) == 3 + 6 * m
        assert _lib.num_params(_lib.BERNOULLI_GAMMA, m) == 5 + 10 * m
    assert lib.attrici_num_params(9, 4) == -1
    assert lib.attrici_num_params(0, 5) == -1
    assert b"bad family" in lib.attrici_last_error()


def test_workspace_query_and_argument_errors():
    lib = _lib.load()
    n = _lib.workspace_bytes(_lib.NORMAL, 43464, 10000)
    assert 50e6 < n < 2.0e9
    # per-phase buffers of the folded passes: Normal 2 x [1461][3][C] float64 (phase-major, fold.cu) + 2 x [C][3][1464]
    # float64 and [C][1464] uint8 (cell-major, fit_fused.cu); other families [1461][11][C] float32
    n_gamma = _lib.workspace_bytes(_lib.GAMMA, 43464, 10000)
    normal_only = (2 * 3 * 8 * 1461 + 2 * 3 * 8 * 1464 + 1464) * 10016
    assert n - normal_only > 50e6 and n_gamma - 4 * 11 * 1461 * 10000 > 50e6
    assert abs((n - normal_only) - (n_gamma - 4 * 11 * 1461 * 10016)) < 8192
    # the Bernoulli-Gamma family also holds the [T x C] MT19937 uniforms
    assert _lib.workspace_bytes(_lib.BERNOULLI_GAMMA, 43464, 10000) > n_gamma + 8 * 43464 * 10000 - 1
    with pytest.raises(_lib.AttriciCudaError, match="T >= 2"):
        _lib.workspace_bytes(_lib.NORMAL, 1, 10)
    var = create_variable("tas").descriptor(4)
    # null pointers and bad shapes are rejected before any CUDA call
    rc = lib.attrici_fit_batch(None, 0, C.byref(var), None, 10, 100, 10, 0, 100, None, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.attrici_last_error()
    rc = lib.attrici_prep_scale_mask(None, 0, C.byref(var), None, None, 5, 100, 10, 0, 100, None, None)
    assert rc == -1 and b"leading dimension" in lib.attrici_last_error()
    rc = lib.attrici_quantile_map(None, 0, None, None, None, 10, 100, 10, None, None, None, None, None, None, 10, None)
    assert rc == -1 and b"descriptor" in lib.attrici_last_error()
    bad = create_variable("tas").descriptor(4)
    bad.modes = 7
    rc = lib.attrici_nll_grad(None, 0, C.byref(bad), None, 10, 100, 10, 0, 100, 1, None, None, None, None, None)
    assert rc == -1 and b"modes" in lib.attrici_last_error()


def test_variable_table_matches_reference_model_table():
    # attrici/variables.py create_variable (:801-834) and SURVEY.md section 8a
    fam = {"tas": _lib.NORMAL, "ps": _lib.NORMAL, "rlds": _lib.NORMAL, "rsds": _lib.NORMAL, "tasskew": _lib.NORMAL,
           "pr": _lib.BERNOULLI_GAMMA, "tasrange": _lib.GAMMA, "hurs": _lib.BETA, "sfcWind": _lib.WEIBULL,
           "wind": _lib.WEIBULL}
    assert set(VARIABLES) == set(fam)
    for k, f in fam.items():
        assert VARIABLES[k].family == f
    with pytest.raises(ValueError, match="not supported"):
        create_variable("huss")
    assert create_variable("pr").parameters == ("p", "mu", "nu")
    assert create_variable("sfcWind").parameters == ("alpha", "beta")


def test_solver_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from attrici_b200.solver import CellBatchSolver

    with pytest.raises(_lib.AttriciCudaError, match="no CPU fallback"):
        CellBatchSolver("tas")
