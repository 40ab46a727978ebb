"""ctypes loader for tests/hostmath (CPU build of the device mathematics; TEST SCAFFOLDING ONLY).

Builds ``tests/hostmath/build/libhostmath.so`` with g++ on first use.  Nothing in the product
package imports this module.
"""

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostmath", "hostmath.cpp")
OUT = os.path.join(HERE, "hostmath", "build", "libhostmath.so")
CSRC = os.path.join(os.path.dirname(HERE), "attrici_b200", "csrc")

NP_CANON = 27
NA, NB, MAXM = 18, 9, 4
TERMS = {"normal": 0, "gamma": 1, "beta": 2, "weibull": 3, "bernoulli": 4}
FAMILIES = {"normal": 0, "gamma": 1, "beta": 2, "weibull": 3, "bernoulli_gamma": 4}
SCALINGS = {"unity": 0, "none": 1, "percent": 2, "range": 3, "pr": 4}
POSTS = {None: 0, "le0_nan": 1, "outside01_nan": 2}

_lib = None


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def lib():
    global _lib
    if _lib is None:
        if _stale():
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", OUT, SRC])
        _lib = C.CDLL(OUT)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def unary(which, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib().hm_unary(C.c_int(which), _p(x, C.c_double), _p(out, C.c_double), C.c_int(x.size))
    return out


def binary(which, a, x):
    a = np.ascontiguousarray(a, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib().hm_binary(C.c_int(which), _p(a, C.c_double), _p(x, C.c_double), _p(out, C.c_double), C.c_int(x.size))
    return out


def gamma_inv_from(a, x0, p):
    """``gamma_p_inv_from`` started at ``x0`` with P, Q and the prefactor of one evaluation there (qmap_pr.cu)."""
    a, x0, p = (np.ascontiguousarray(v, dtype=np.float64) for v in (a, x0, p))
    out = np.empty_like(p)
    lib().hm_gamma_inv_from(_p(a, C.c_double), _p(x0, C.c_double), _p(p, C.c_double), _p(out, C.c_double), C.c_int(p.size))
    return out


def ternary(which, a, b, x):
    a, b, x = (np.ascontiguousarray(v, dtype=np.float64) for v in (a, b, x))
    out = np.empty_like(x)
    lib().hm_ternary(C.c_int(which), _p(a, C.c_double), _p(b, C.c_double), _p(x, C.c_double), _p(out, C.c_double),
                     C.c_int(x.size))
    return out


def family(kind, y, ea, eb, f32=False):
    y, ea, eb = (np.ascontiguousarray(v, dtype=np.float64) for v in (y, ea, eb))
    out = np.empty((y.size, 6))
    fn = lib().hm_family_f32 if f32 else lib().hm_family_f64
    fn(C.c_int(TERMS[kind]), _p(y, C.c_double), _p(ea, C.c_double), _p(eb, C.c_double), _p(out, C.c_double),
       C.c_int(y.size))
    return out


def canon_index(family_name, modes):
    """(slot, canonical index) of every reference-layout coefficient."""
    fam = FAMILIES[family_name]
    n = lib().hm_num_params(C.c_int(fam), C.c_int(modes))
    out = []
    for j in range(n):
        s, c = C.c_int(), C.c_int()
        lib().hm_ref_to_canon(C.c_int(fam), C.c_int(modes), C.c_int(j), C.byref(s), C.byref(c))
        out.append((s.value, c.value))
    return out


def nll_grad_cell(term, modes, ys, days, gmt, i0, i1, theta_canon, fp64=True, n_chunks=8, want_fisher=True):
    ys = np.ascontiguousarray(ys, dtype=np.float32)
    days = np.ascontiguousarray(days, dtype=np.int32)
    gmt = np.ascontiguousarray(gmt, dtype=np.float64)
    th = np.ascontiguousarray(theta_canon, dtype=np.float64)
    nll = C.c_double()
    grad = np.empty(NP_CANON)
    fisher = np.empty((NP_CANON, NP_CANON))
    lib().hm_nll_grad_cell(C.c_int(TERMS[term]), C.c_int(modes), C.c_int(len(ys)), C.c_int(i0), C.c_int(i1),
                           _p(ys, C.c_float), _p(days, C.c_int32), _p(gmt, C.c_double), C.c_int(int(fp64)),
                           C.c_int(n_chunks), _p(th, C.c_double), C.byref(nll), _p(grad, C.c_double),
                           _p(fisher, C.c_double) if want_fisher else None)
    return nll.value, grad, fisher


def fit_cell(term, modes, ys, days, gmt, i0, i1, fp64=False, zero_init=False, max_pass=60, tol_pred=1e-11,
             lambda0=1e-3, n_chunks=8, want_trace=False):
    ys = np.ascontiguousarray(ys, dtype=np.float32)
    days = np.ascontiguousarray(days, dtype=np.int32)
    gmt = np.ascontiguousarray(gmt, dtype=np.float64)
    th = np.empty(NP_CANON)
    logp, status, npass = C.c_double(), C.c_int(), C.c_int()
    trace = np.zeros((max_pass, 4))
    lib().hm_fit_cell(C.c_int(TERMS[term]), C.c_int(modes), C.c_int(len(ys)), C.c_int(i0), C.c_int(i1),
                      _p(ys, C.c_float), _p(days, C.c_int32), _p(gmt, C.c_double), C.c_int(int(fp64)),
                      C.c_int(int(zero_init)), C.c_int(max_pass), C.c_double(tol_pred), C.c_double(lambda0),
                      C.c_int(n_chunks), _p(th, C.c_double), C.byref(logp), C.byref(status), C.byref(npass),
                      _p(trace, C.c_double) if want_trace else None)
    out = {"theta": th, "logp": logp.value, "status": status.value, "n_pass": npass.value}
    if want_trace:
        out["trace"] = trace[: npass.value]  # per pass: day-term nll of the trial, accepted f, predicted decrease, lambda
    return out


def qmap_cell(family_name, scaling, post, modes, y, ys, days, gmt, params, datamin, scale, uniforms=None):
    y = np.ascontiguousarray(y, dtype=np.float32)
    ys = np.ascontiguousarray(ys, dtype=np.float32)
    days = np.ascontiguousarray(days, dtype=np.int32)
    gmt = np.ascontiguousarray(gmt, dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    T = len(y)
    cfact, cs = np.empty(T), np.empty(T)
    rep = np.empty(T, dtype=np.int32)
    pr, pc = np.empty((T, 3)), np.empty((T, 3))
    u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
    lib().hm_qmap_cell(C.c_int(FAMILIES[family_name]), C.c_int(SCALINGS[scaling]), C.c_int(POSTS[post]),
                       C.c_int(modes), C.c_int(T), _p(y, C.c_float), _p(ys, C.c_float), _p(days, C.c_int32),
                       _p(gmt, C.c_double), _p(params, C.c_double), C.c_double(datamin), C.c_double(scale),
                       _p(u, C.c_double) if u is not None else None, _p(cfact, C.c_double), _p(cs, C.c_double),
                       _p(rep, C.c_int32), _p(pr, C.c_double), _p(pc, C.c_double))
    return {"cfact": cfact, "cfact_scaled": cs, "replaced": rep, "par_ref": pr, "par_cf": pc}


def mt19937(seed, n):
    out = np.empty(n)
    lib().hm_mt19937(C.c_uint32(seed), C.c_int(n), _p(out, C.c_double))
    return out
