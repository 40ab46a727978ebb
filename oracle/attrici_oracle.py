"""CPU oracle for ATTRICI's per-gridcell detrending path (TEST INFRASTRUCTURE ONLY).

This file is a NumPy/SciPy restatement of the reference algorithm, written to CHECK the
CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product package
``attrici_b200`` never does (it fails loudly when its CUDA library is missing).

Parity pinning (see DESIGN.md "Oracle"): ``tests/golden/make_golden.py`` runs the
UNMODIFIED reference (``/root/reference/attrici``; xarray/func_timeout/tomlkit replaced by
the import stand-ins under ``oracle/standins``) on the reference's own fixtures and stores
its traces / counterfactuals in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function below against those files and against the reference's golden
``.h5`` targets (reference ``tests/test_detrend_pymc.py:63-160``).

Reference lines restated (relative to /root/reference):
  attrici/util.py:102-104                          -> oscillations()
  attrici/detrend.py:698-707                       -> gmt_on_time_axis()
  attrici/detrend.py:285                           -> cell_seed()
  attrici/detrend.py:311, 392-411                  -> detrend_cell() (inf->nan, replacement)
  attrici/variables.py:40-68, 92-131, 363-403,
      574-591, 619-623, 722-732, 770-780           -> scale_variable(), rescale()
  attrici/variables.py:287-303, 422-480, 645-653,
      691-698                                      -> quantile_map()
  attrici/distributions.py:81-216                  -> cdf(), invcdf()
  attrici/estimation/model_scipy.py:104-178,
      225-278, 311-373, 410-503                    -> Problem.logpost() (+ analytic derivatives,
                                                      which the reference does not have)
  attrici/estimation/model_scipy.py:519-527        -> fit_lbfgsb()
  attrici/estimation/model_scipy.py:547-570        -> estimate_params()

Documented deviation ("link correction", SURVEY.md section 8c defect 1): the reference's
``PredictorDependentParam.estimate`` (model_scipy.py:173-178) returns the linear predictor
WITHOUT applying the link, which makes the objective NaN at theta=0 for pr / tasrange / hurs /
sfcWind.  Like PyMC5 (model_pymc5.py:253-256) this oracle applies the link
(``apply_dependent_link=True``); with ``False`` it reproduces the shipped behaviour.
For identity-link variables (tas, ps, rlds, rsds, tasskew) the two are identical.
"""

from dataclasses import dataclass, field

import numpy as np
from scipy import optimize, special, stats

LN_SQRT_2PI = 0.5 * np.log(2.0 * np.pi)
PR_THRESHOLD = 0.0000011574  # variables.py:350

# --------------------------------------------------------------------------------------
# model table (variables.py:306-834): family, parameter order (= theta layout), links
# --------------------------------------------------------------------------------------
# each parameter: (name, link, dependent-on-GMT)
FAMILY_PARAMS = {
    "normal": (("mu", "identity", True), ("sigma", "exp", False)),
    "gamma": (("mu", "exp", True), ("nu", "exp", False)),
    "beta": (("mu", "invlogit", True), ("phi", "exp", False)),
    "weibull": (("alpha", "exp", False), ("beta", "exp", True)),
    "bernoulli_gamma": (("p", "invlogit", True), ("mu", "exp", True), ("nu", "exp", False)),
}

# variable -> (family, scaling kind, lower mask thr, upper mask thr, qmap post rule)
VARIABLES = {
    "tas": ("normal", "unity", None, None, None),
    "ps": ("normal", "unity", None, None, None),
    "rlds": ("normal", "unity", None, None, None),
    "rsds": ("normal", "unity", None, None, "le0_nan"),
    "tasskew": ("normal", "none", 0.0001, 0.9999, "outside01_nan"),
    "pr": ("bernoulli_gamma", "pr", None, None, None),
    "tasrange": ("gamma", "range", 0.01, None, None),
    "hurs": ("beta", "percent", 0.01, 99.99, None),
    "sfcWind": ("weibull", "range", 0.01, None, None),
    "wind": ("weibull", "range", 0.01, None, None),
}

LINKS = {
    "identity": lambda x: x,
    "exp": np.exp,
    "invlogit": lambda x: 1.0 / (1.0 + np.exp(-x)),
}


def n_params(family, modes):
    return sum((2 + 4 * modes) if dep else (1 + 2 * modes) for _, _, dep in FAMILY_PARAMS[family])


def param_slices(family, modes):
    """theta layout (model_scipy.py:410-421): blocks in dict order."""
    out, i = {}, 0
    for name, _, dep in FAMILY_PARAMS[family]:
        n = (2 + 4 * modes) if dep else (1 + 2 * modes)
        out[name] = slice(i, i + n)
        i += n
    return out


def prior_sigma(dep, modes):
    """Per-coefficient prior std (np.inf = no prior), model_scipy.py:142-171, 256-270.

    Only the first ``modes`` entries of each ``2*modes`` Fourier block (the cos
    coefficients) carry a prior.
    """
    m = modes
    cos_s = [1.0 / (2 * i + 1) for i in range(m)]
    if dep:
        return np.array([1.0, 0.1] + cos_s + [np.inf] * m + [0.1] * m + [np.inf] * m)
    return np.array([1.0] + cos_s + [np.inf] * m)


# --------------------------------------------------------------------------------------
# predictor / basis
# --------------------------------------------------------------------------------------
def gmt_on_time_axis(gmt, n_time):
    """detrend.py:698-707 (default, not ``full_extrapolation``): stretch the GMT series
    over the normalised observation index and min-max scale it to [0, 1].

    ``t_scaled = (time - time.min()) / (time.max() - time.min())`` for a gap-free daily
    axis equals ``i / (n_time - 1)`` computed in float64.
    """
    days = np.arange(n_time, dtype=np.int64)
    t_scaled = days.astype(np.float64) / np.float64(n_time - 1)
    g = np.interp(t_scaled, np.linspace(0, 1, len(gmt)), gmt)
    return (g - g.min()) / (g.max() - g.min())


def oscillations(days, modes):
    """util.py:102-104 with ``t`` given as integer days; phase origin = ``days.min()``.

    ``(t - t.min()) / (365 d + 6 h)`` on datetime64[ns] is an exact-integer float division,
    bit-identical to ``(days - days.min()) / 365.25`` in float64.
    """
    days = np.asarray(days)
    t_scaled = (days - days.min()).astype(np.float64) / 365.25
    x = (2 * np.pi * (np.arange(modes) + 1)) * t_scaled[:, None]
    return np.concatenate((np.cos(x), np.sin(x)), axis=1)


def design(days, gmt, modes, dep):
    """Full design matrix of one parameter block in theta order.

    dependent  (model_scipy.py:129-140, 173-197): [1, T, cos, sin, T*cos, T*sin]
    independent(model_scipy.py:250-255, 272-289): [1, cos, sin]
    """
    osc = oscillations(days, modes)
    one = np.ones((len(days), 1))
    if dep:
        g = np.asarray(gmt, dtype=np.float64)[:, None]
        return np.concatenate([one, g, osc, g * osc], axis=1)
    return np.concatenate([one, osc], axis=1)


# --------------------------------------------------------------------------------------
# scaling (variables.py)
# --------------------------------------------------------------------------------------
def gamma_fit_scale(x):
    """Pr.scale normaliser, variables.py:388-393: ``fscale * fa**0.5`` of
    ``scipy.stats.gamma.fit(x, floc=0)`` (dtype follows the float32 input)."""
    fa, _, fscale = stats.gamma.fit(x, floc=0)
    return fscale * fa**0.5


def scale_variable(variable, y):
    """Return (y_scaled float32 with NaN on masked days, scaling dict, y_out).

    ``y_out`` is the caller-visible ``data`` after the constructor ran: hurs / tasrange /
    sfcWind mask IN PLACE (variables.py:583, 723, 771), so masked days are NaN in the
    output ``y`` and in the replacement source as well.  All arithmetic is float32, as
    xarray keeps the file dtype.
    """
    family, kind, lo, hi, _ = VARIABLES[variable]
    y = np.array(y, dtype=np.float32, copy=True)
    y[np.isinf(y)] = np.nan  # detrend.py:311
    scaling = {}
    if kind == "unity":  # variables.py:92-111
        datamin = np.nanmin(y)
        scale = np.nanmax(y) - datamin
        ys = (y - datamin) / scale
        scaling = {"datamin": float(datamin), "scale": float(scale)}
        return ys, scaling, y
    if kind == "none":  # Tasskew, variables.py:619-623 (masks a copy)
        ys = y.copy()
        ys[ys <= np.float32(lo)] = np.nan
        ys[ys >= np.float32(hi)] = np.nan
        return ys, scaling, y
    if kind == "percent":  # Hurs, variables.py:574-591
        y[y <= np.float32(lo)] = np.nan
        y[y >= np.float32(hi)] = np.nan
        ys = y / np.float32(100.0)
        return ys, {"scale": 100.0}, y
    if kind == "range":  # Tasrange / Wind, variables.py:722-732, 770-780
        y[y <= np.float32(lo)] = np.nan
        scale = np.nanmax(y) - np.nanmin(y)
        ys = y / scale
        return ys, {"scale": float(scale)}, y
    if kind == "pr":  # variables.py:363-403
        ys = y - np.float32(PR_THRESHOLD)
        ys[ys <= 0] = np.nan
        if not np.any(~np.isnan(ys)):
            return ys, {"scale": np.nan}, y
        scale = gamma_fit_scale(ys[~np.isnan(ys)])
        ys = ys / scale
        return ys.astype(np.float32), {"scale": float(scale)}, y
    raise ValueError(kind)


def rescale(variable, cfact_scaled, scaling):
    """variables.py:71-89, 114-131, 483-486, 657."""
    kind = VARIABLES[variable][1]
    if kind == "unity":
        return cfact_scaled * scaling["scale"] + scaling["datamin"]
    if kind == "none":
        return cfact_scaled * 1.0
    if kind in ("percent", "range"):
        return cfact_scaled * scaling["scale"]
    if kind == "pr":
        data = cfact_scaled * scaling["scale"] + PR_THRESHOLD
        data[cfact_scaled <= 0] = 0
        return data
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# per-day log-likelihood and its derivatives w.r.t. the linear predictors
# --------------------------------------------------------------------------------------
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def family_terms(kind, y, eta_a, eta_b):
    """Per-day NLL and derivatives for one likelihood term.

    ``kind`` in normal|gamma|beta|weibull|bernoulli.  ``eta_a`` is the linear predictor of
    the GMT-dependent parameter, ``eta_b`` of the independent one (None for bernoulli).
    Returns (nll, r_a, r_b, w_aa, w_ab, w_bb) with r = d nll/d eta, w = d2 nll/d eta2
    (formulas: SURVEY.md appendix B, re-derived from the scipy logpdf definitions).
    """
    if kind == "normal":
        e = np.exp(-eta_b)
        z = (y - eta_a) * e
        nll = 0.5 * z * z + eta_b + LN_SQRT_2PI
        return nll, -z * e, 1.0 - z * z, e * e, 2.0 * z * e, 2.0 * z * z
    if kind == "gamma":
        a = np.exp(2.0 * eta_b)
        u = y * np.exp(-eta_a)
        lnu = np.log(y) - eta_a
        lna = 2.0 * eta_b
        d = lnu + lna + 1.0 - u - special.digamma(a)
        nll = -((a - 1.0) * np.log(y) - a * u - special.gammaln(a) - a * eta_a + a * lna)
        return (
            nll,
            -a * (u - 1.0),
            -2.0 * a * d,
            a * u,
            2.0 * a * (1.0 - u),
            4.0 * a * a * special.polygamma(1, a) - 4.0 * a - 4.0 * a * d,
        )
    if kind == "weibull":
        c = np.exp(eta_b)
        L = np.log(y) - eta_a
        r = np.exp(c * L)
        nll = -(eta_b + (c - 1.0) * L - r - eta_a)
        return (
            nll,
            -c * (r - 1.0),
            -(1.0 + c * L * (1.0 - r)),
            c * c * r,
            -c * (r - 1.0) - c * c * L * r,
            -c * L * (1.0 - r) + c * c * L * L * r,
        )
    if kind == "beta":
        mu = _sigmoid(eta_a)
        phi = np.exp(eta_b)
        A, B = mu * phi, (1.0 - mu) * phi
        ly, l1y = np.log(y), np.log1p(-y)
        psiA, psiB, psiP = special.digamma(A), special.digamma(B), special.digamma(phi)
        tA, tB, tP = special.polygamma(1, A), special.polygamma(1, B), special.polygamma(1, phi)
        m1 = mu * (1.0 - mu)
        delta = ly - l1y - psiA + psiB
        ell = (
            (A - 1.0) * ly + (B - 1.0) * l1y - special.gammaln(A) - special.gammaln(B)
            + special.gammaln(phi)
        )
        dl_a = phi * m1 * delta
        dl_b = A * (ly - psiA) + B * (l1y - psiB) + phi * psiP
        d2_aa = phi * m1 * (1.0 - 2.0 * mu) * delta - (phi * m1) ** 2 * (tA + tB)
        d2_ab = phi * m1 * (delta - A * tA + B * tB)
        d2_bb = dl_b - A * A * tA - B * B * tB + phi * phi * tP
        return -ell, -dl_a, -dl_b, -d2_aa, -d2_ab, -d2_bb
    if kind == "bernoulli":
        p = _sigmoid(eta_a)
        # y = 1 on a dry day (model_scipy.py:444)
        nll = -(special.xlogy(y, p) + special.xlog1py(1.0 - y, -p))
        return nll, -(y - p), None, p * (1.0 - p), None, None
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# the fit problem of one cell
# --------------------------------------------------------------------------------------
@dataclass
class Term:
    """One likelihood term (model_scipy.py:433-498): its data and design matrices."""

    kind: str
    y: np.ndarray  # float64 observations (dry indicator for bernoulli)
    name_a: str
    x_a: np.ndarray  # [n, 2+4m]
    name_b: str = None
    x_b: np.ndarray = None  # [n, 1+2m]


@dataclass
class Problem:
    family: str
    modes: int
    terms: list
    slices: dict
    n: int
    prec: np.ndarray = field(default=None)  # prior precision (0 = flat)
    prior_const: float = 0.0
    apply_dependent_link: bool = True

    # ---- the reference objective, restated with the same scipy calls ---------------
    def logpost_scipy(self, theta):
        """model_scipy.py:311-331 summed over distributions (:520), same scipy.stats
        calls and the same per-coefficient scalar prior calls as the reference, so that
        timing it is representative of the reference's cost per evaluation."""
        theta = np.asarray(theta, dtype=np.float64)
        m = self.modes
        res = 0.0
        for t in self.terms:
            vals = {}
            for name, x in ((t.name_a, t.x_a), (t.name_b, t.x_b)):
                if name is None:
                    continue
                link, dep = next((l, d) for n_, l, d in FAMILY_PARAMS[self.family] if n_ == name)
                th = theta[self.slices[name]]
                if dep:
                    lp = stats.norm.logpdf(th[0], loc=0, scale=1)
                    lp += stats.norm.logpdf(th[1], loc=0, scale=0.1)
                    lp += np.sum([stats.norm.logpdf(th[2 + i], loc=0, scale=1 / (2 * i + 1)) for i in range(m)])
                    lp += np.sum([stats.norm.logpdf(th[2 + 2 * m + i], loc=0, scale=0.1) for i in range(m)])
                    eta = np.dot(x[:, 2:], th[2:]) + th[0] + th[1] * x[:, 1]
                    vals[name] = LINKS[link](eta) if self.apply_dependent_link else eta
                else:
                    lp = stats.norm.logpdf(th[0], loc=0, scale=1)
                    lp += np.sum([stats.norm.logpdf(th[1 + i], loc=0, scale=1 / (2 * i + 1)) for i in range(m)])
                    vals[name] = LINKS[link](np.dot(x[:, 1:], th[1:]) + th[0])
                res += lp
            if t.kind == "normal":
                ll = stats.norm.logpdf(t.y, loc=vals["mu"], scale=vals["sigma"])
            elif t.kind == "gamma":
                ll = stats.gamma.logpdf(t.y, vals["nu"] ** 2, scale=vals["mu"] / vals["nu"] ** 2)
            elif t.kind == "beta":
                ll = stats.beta.logpdf(t.y, vals["mu"] * vals["phi"], (1 - vals["mu"]) * vals["phi"])
            elif t.kind == "weibull":
                ll = stats.weibull_min.logpdf(t.y, c=vals["alpha"], scale=vals["beta"])
            elif t.kind == "bernoulli":
                ll = stats.bernoulli.logpmf(t.y.astype(int), vals["p"])
            res += np.sum(ll)
        return res

    # ---- closed-form objective with analytic gradient / Hessian --------------------
    def nll(self, theta, order=0):
        """Negative log-posterior; order 1 adds the gradient, 2 the exact Hessian."""
        theta = np.asarray(theta, dtype=np.float64)
        f = 0.5 * np.sum(self.prec * theta * theta) - self.prior_const
        g = self.prec * theta if order >= 1 else None
        H = np.diag(self.prec.copy()) if order >= 2 else None
        for t in self.terms:
            sa = self.slices[t.name_a]
            eta_a = t.x_a @ theta[sa]
            if t.name_b is not None:
                sb = self.slices[t.name_b]
                eta_b = t.x_b @ theta[sb]
            else:
                eta_b = None
            with np.errstate(all="ignore"):
                nll, ra, rb, waa, wab, wbb = family_terms(t.kind, t.y, eta_a, eta_b)
            f += np.sum(nll)
            if order >= 1:
                g[sa] += t.x_a.T @ ra
                if rb is not None:
                    g[sb] += t.x_b.T @ rb
            if order >= 2:
                H[sa, sa] += t.x_a.T @ (waa[:, None] * t.x_a)
                if rb is not None:
                    hab = t.x_a.T @ (wab[:, None] * t.x_b)
                    H[sa, sb] += hab
                    H[sb, sa] += hab.T
                    H[sb, sb] += t.x_b.T @ (wbb[:, None] * t.x_b)
        if order == 0:
            return f
        if order == 1:
            return f, g
        return f, g, H


def build_problem(family, y_scaled, gmt, days, modes, fit_slice=None, apply_dependent_link=True):
    """Variable.create_model (variables.py:323-334, 406-419, ...) + ModelScipy.__init__
    (model_scipy.py:379-503) for one cell.

    ``fit_slice`` selects ``subset_times`` (detrend.py:720-722).  Non-pr variables drop
    NaN days (``only_valid``); pr keeps every day for ``p`` and the wet days for mu/nu
    (model_scipy.py:424-431).  Each block's Fourier phase origin is the first day of the
    time vector it is given (util.py:102).
    """
    y_scaled = np.asarray(y_scaled)
    days = np.asarray(days)
    gmt = np.asarray(gmt, dtype=np.float64)
    if fit_slice is None:
        fit_slice = slice(0, len(days))
    yw, dw, gw = y_scaled[fit_slice], days[fit_slice], gmt[fit_slice]
    valid = ~np.isnan(yw)
    slices = param_slices(family, modes)
    P = n_params(family, modes)
    prec = np.zeros(P)
    const = 0.0
    for name, _, dep in FAMILY_PARAMS[family]:
        s = prior_sigma(dep, modes)
        has = np.isfinite(s)
        prec[slices[name]] = np.where(has, 1.0 / np.where(has, s, 1.0) ** 2, 0.0)
        const += np.sum(-np.log(s[has]) - LN_SQRT_2PI)
    terms = []
    yv = yw[valid].astype(np.float64)
    if family == "bernoulli_gamma":
        terms.append(
            Term("gamma", yv, "mu", design(dw[valid], gw[valid], modes, True), "nu",
                 design(dw[valid], gw[valid], modes, False))
        )
        terms.append(Term("bernoulli", np.isnan(yw).astype(np.float64), "p", design(dw, gw, modes, True)))
    else:
        (na, _, _), (nb, _, _) = sorted(FAMILY_PARAMS[family], key=lambda p: not p[2])
        terms.append(
            Term(family, yv, na, design(dw[valid], gw[valid], modes, True), nb,
                 design(dw[valid], gw[valid], modes, False))
        )
    return Problem(family, modes, terms, slices, P, prec, const, apply_dependent_link)


def fit_lbfgsb(problem):
    """model_scipy.py:519-527: L-BFGS-B from theta=0 with finite-difference gradients."""
    with np.errstate(all="ignore"):
        res = optimize.minimize(lambda p: -problem.logpost_scipy(p), np.zeros(problem.n), method="L-BFGS-B")
    return {"logp": -res.fun, "params": res.x, "nit": res.nit, "nfev": res.nfev}


# Levenberg-Marquardt damped Newton on the exact Hessian: the SAME iteration the CUDA
# solver runs (attrici_b200/csrc/step.cu), kept here in float64 as the "exact MAP" oracle.
LM_DEFAULTS = dict(max_iter=60, lam0=1e-3, lam_up=8.0, lam_down=0.2, lam_min=1e-9, lam_max=1e12,
                   tol_pred=1e-11, accept_slack=1e-9)


def lm_solve(H, g, lam):
    """Solve (H + lam*D) delta = -g, D = max(|diag H|, tiny); raise lam until SPD."""
    d = np.maximum(np.abs(np.diag(H)), 1e-12)
    while True:
        try:
            c = np.linalg.cholesky(H + np.diag(lam * d))
            y = np.linalg.solve(c, -g)
            return np.linalg.solve(c.T, y), lam
        except np.linalg.LinAlgError:
            lam = max(lam * 10.0, 1e-6)
            if lam > 1e15:
                raise


def fit_newton(problem, **kw):
    o = {**LM_DEFAULTS, **kw}
    theta = np.zeros(problem.n)
    f, g, H = problem.nll(theta, 2)
    lam = o["lam0"]
    n_pass, status = 1, 1
    for _ in range(o["max_iter"]):
        delta, lam = lm_solve(H, g, lam)
        pred = -(g @ delta) - 0.5 * delta @ H @ delta
        if pred <= o["tol_pred"] * max(abs(f), 1.0):
            status = 0
            break
        ft, gt, Ht = problem.nll(theta + delta, 2)
        n_pass += 1
        ok = np.isfinite(ft) and np.all(np.isfinite(gt)) and np.all(np.isfinite(Ht))
        if ok and ft <= f + o["accept_slack"] * abs(f) and (f - ft) >= 1e-4 * pred - o["accept_slack"] * abs(f):
            rho = (f - ft) / pred
            theta, f, g, H = theta + delta, ft, gt, Ht
            if rho > 0.75:
                lam = max(lam * o["lam_down"], o["lam_min"])
            elif rho < 0.25:
                lam = lam * 2.0
        else:
            lam = max(lam * o["lam_up"], 1e-4)
            if lam > o["lam_max"]:
                status = 2
                break
    return {"logp": -f, "params": theta, "n_pass": n_pass, "status": status, "grad_inf": np.abs(g).max()}


# --------------------------------------------------------------------------------------
# distribution parameters, cdf / invcdf (model_scipy.py:547-570, distributions.py)
# --------------------------------------------------------------------------------------
def estimate_params(family, theta, gmt, days, modes, apply_dependent_link=True):
    """Per-day distribution parameters for the FULL series (phase origin = days.min())."""
    out = {}
    sl = param_slices(family, modes)
    for name, link, dep in FAMILY_PARAMS[family]:
        eta = design(days, gmt, modes, dep) @ np.asarray(theta)[sl[name]]
        out[name] = LINKS[link](eta) if (apply_dependent_link or not dep) else eta
    return out


def cdf(family, p, y):
    if family == "normal":
        return stats.norm.cdf(y, loc=p["mu"], scale=p["sigma"])
    if family == "gamma":
        return stats.gamma.cdf(y, p["nu"] ** 2.0, scale=p["mu"] / p["nu"] ** 2.0)
    if family == "bernoulli_gamma":
        return p["p"] + (1 - p["p"]) * stats.gamma.cdf(y, p["nu"] ** 2.0, scale=p["mu"] / p["nu"] ** 2.0)
    if family == "beta":
        return stats.beta.cdf(y, p["mu"] * p["phi"], (1 - p["mu"]) * p["phi"])
    if family == "weibull":
        return stats.weibull_min.cdf(y, p["alpha"], scale=p["beta"])
    raise ValueError(family)


def invcdf(family, p, q):
    if family == "normal":
        return stats.norm.ppf(q, loc=p["mu"], scale=p["sigma"])
    if family == "gamma":
        return stats.gamma.ppf(q, p["nu"] ** 2.0, scale=p["mu"] / p["nu"] ** 2.0)
    if family == "bernoulli_gamma":
        return stats.gamma.ppf((q - p["p"]) / (1 - p["p"]), p["nu"] ** 2.0, scale=p["mu"] / p["nu"] ** 2.0)
    if family == "beta":
        return stats.beta.ppf(q, p["mu"] * p["phi"], (1 - p["mu"]) * p["phi"])
    if family == "weibull":
        return stats.weibull_min.ppf(q, p["alpha"], scale=p["beta"])
    raise ValueError(family)


def cell_seed(seed, lat, lon):
    """detrend.py:285."""
    return (seed + int(lat * 1e9) + int(lon * 1e6)) % 2**32


def quantile_map(variable, y_scaled, ref, cf, uniforms=None):
    """variables.py:287-303 and the Pr / Rsds / Tasskew overrides.

    For pr, ``uniforms`` are the ``np.random.rand(len(y))`` draws of variables.py:458
    (the caller seeds the legacy global RNG as detrend.py:285 does).
    """
    family, _, _, _, post = VARIABLES[variable]
    with np.errstate(all="ignore"):
        if family != "bernoulli_gamma":
            res = invcdf(family, cf, cdf(family, ref, y_scaled))
            if post == "le0_nan":
                res[res <= 0] = np.nan
            elif post == "outside01_nan":
                res[res >= 1] = np.nan
                res[res <= 0] = np.nan
            return res
        y = np.array(y_scaled, copy=True)
        dry = np.isnan(y)
        y[dry] = 0
        q = cdf(family, ref, y)
        drier_cf = cf["p"] > ref["p"]
        qm0 = drier_cf & (q > cf["p"])
        inv = invcdf(family, cf, q)
        out = np.zeros(len(y))
        out[qm0] = inv[qm0]
        qm1 = ~drier_cf & ~dry
        out[qm1] = inv[qm1]
        rnd = uniforms * ref["p"]
        to_wet = ~qm1 & (rnd > cf["p"])
        out[to_wet] = inv[to_wet]
        return out


def replace_invalid(cfact, y_out):
    """detrend.py:397-411: NaN / +Inf / -Inf from the mapping fall back to the original."""
    cfact = np.array(cfact, dtype=np.float64, copy=True)
    replaced = np.zeros_like(cfact)
    for test, code in ((np.isnan, np.nan), (lambda c: c == np.inf, np.inf), (lambda c: c == -np.inf, -np.inf)):
        idx = test(cfact)
        cfact[idx] = y_out[idx]
        replaced[idx] = code
    return cfact, replaced


def detrend_cell(variable, y, gmt_scaled, days, modes=4, fit_stop=None, fit_start=0, seed=0, lat=0.0,
                 lon=0.0, solver="newton", apply_dependent_link=True, trace=None):
    """fit_and_detrend_cell (detrend.py:258-416) for one cell, NetCDF I/O excluded."""
    family = VARIABLES[variable][0]
    np.random.seed(cell_seed(seed, lat, lon))
    ys, scaling, y_out = scale_variable(variable, y)
    n = len(ys)
    fit_slice = slice(fit_start, n if fit_stop is None else fit_stop)
    prob = build_problem(family, ys, gmt_scaled, days, modes, fit_slice, apply_dependent_link)
    if trace is None:
        trace = fit_lbfgsb(prob) if solver == "lbfgsb" else fit_newton(prob)
    ref = estimate_params(family, trace["params"], gmt_scaled, days, modes, apply_dependent_link)
    cf = estimate_params(family, trace["params"], np.zeros_like(gmt_scaled), days, modes, apply_dependent_link)
    uniforms = np.random.rand(n) if family == "bernoulli_gamma" else None
    cfs = quantile_map(variable, ys, ref, cf, uniforms)
    cfact = rescale(variable, np.asarray(cfs, dtype=np.float64), scaling)
    cfact, replaced = replace_invalid(cfact, y_out)
    return {"y_scaled": ys, "scaling": scaling, "trace": trace, "ref": ref, "cf": cf, "cfact_scaled": cfs,
            "cfact": cfact, "replaced": replaced, "logp": trace["logp"], "y_out": y_out}
