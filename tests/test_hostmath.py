"""CPU unit tests of the DEVICE mathematics: attrici_b200/csrc/*.cuh compiled with g++ (tests/hostmath)
and compared with scipy.special, the oracle and the golden fixtures.  This validates formulas and control
flow without a GPU; the CUDA kernels themselves are covered by the `-m gpu` tests through the C ABI."""

import numpy as np
import pytest
from scipy import special

import helpers
import hostmath_lib as hm
from oracle import attrici_oracle as oracle


def test_digamma_trigamma():
    x = np.concatenate([np.logspace(-3, 3, 400), np.linspace(0.5, 30, 300)])
    np.testing.assert_allclose(hm.unary(0, x), special.digamma(x), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(hm.unary(1, x), special.polygamma(1, x), rtol=1e-12)
    x32 = x[x > 0.05]
    np.testing.assert_allclose(hm.unary(4, x32), special.digamma(x32), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(hm.unary(5, x32), special.polygamma(1, x32), rtol=2e-5)


def test_normal_cdf_ppf():
    z = np.linspace(-36, 9, 4001)
    np.testing.assert_allclose(hm.unary(2, z), special.ndtr(z), rtol=2e-13, atol=1e-300)
    p = np.concatenate([np.logspace(-300, -1, 500), np.linspace(0.1, 0.9, 200), 1 - np.logspace(-16, -1, 200), [0.0, 1.0]])
    np.testing.assert_allclose(hm.unary(3, p), special.ndtri(p), rtol=2e-14, atol=1e-15)
    assert np.isnan(hm.unary(3, np.array([-0.1, 1.1, np.nan]))).all()


def test_incomplete_gamma():
    rng = np.random.default_rng(0)
    a = np.exp(rng.uniform(np.log(0.05), np.log(500), 4000))
    x = a * np.exp(rng.normal(0, 1.2, 4000))
    np.testing.assert_allclose(hm.binary(0, a, x), special.gammainc(a, x), rtol=3e-12, atol=1e-300)
    np.testing.assert_allclose(hm.binary(1, a, x), special.gammaincc(a, x), rtol=3e-12, atol=1e-300)
    p = rng.uniform(0, 1, 4000)
    p[:200] = np.logspace(-12, -3, 200)
    p[200:400] = 1 - np.logspace(-12, -3, 200)
    got, want = hm.binary(2, a, p), special.gammaincinv(a, p)
    np.testing.assert_allclose(got, want, rtol=1e-10)
    assert hm.binary(2, np.array([2.0]), np.array([0.0]))[0] == 0.0
    assert np.isinf(hm.binary(2, np.array([2.0]), np.array([1.0]))[0])


def test_gamma_inverse_from_the_factual_point():
    """The inverse of the precipitation quantile map (qmap_pr.cu): started at the factual point x0 with P(a, x0),
    Q(a, x0) known, target quantile Q_cf = beta * Q_ref (the dry-day probabilities of the two distributions differ by a
    few percent; wider here), Halley steps ended by the cubic rule at 3e-4 x - against SciPy's gammaincinv.  The rule
    leaves a small multiple of (3e-4)^3 relative; 1e-9 is asked (the path's tolerance on the counterfactual is 1e-4)."""
    rng = np.random.default_rng(3)
    n = 20000
    a = np.exp(rng.uniform(np.log(0.15), np.log(30), n))
    x0 = rng.gamma(a)
    x0 = np.maximum(x0, 1e-12)
    beta = np.exp(rng.uniform(np.log(0.5), np.log(1.5), n))
    q_ref = special.gammaincc(a, x0)
    ok = (q_ref > 1e-12) & (beta * q_ref < 1 - 1e-9)
    a, x0, beta, q_ref = a[ok], x0[ok], beta[ok], q_ref[ok]
    p = 1.0 - beta * q_ref
    got = hm.gamma_inv_from(a, x0, p)
    want = special.gammaincinv(a, p)
    good = want > 0
    np.testing.assert_allclose(got[good], want[good], rtol=1e-9)
    # equal probabilities: the starting point is the root, no iteration moves it beyond rounding
    same = hm.gamma_inv_from(a, x0, special.gammainc(a, x0))
    np.testing.assert_allclose(same, x0, rtol=1e-10)


def test_incomplete_beta():
    rng = np.random.default_rng(1)
    a = np.exp(rng.uniform(np.log(0.2), np.log(300), 4000))
    b = np.exp(rng.uniform(np.log(0.2), np.log(300), 4000))
    x = rng.beta(a, b)
    x = np.clip(x, 1e-12, 1 - 1e-12)
    np.testing.assert_allclose(hm.ternary(0, a, b, x), special.betainc(a, b, x), rtol=5e-12, atol=1e-300)
    p = rng.uniform(1e-6, 1 - 1e-6, 4000)
    got, want = hm.ternary(1, a, b, p), special.betaincinv(a, b, p)
    np.testing.assert_allclose(got, want, rtol=1e-9)


@pytest.mark.parametrize("kind", ["normal", "gamma", "beta", "weibull", "bernoulli"])
def test_family_terms_against_oracle(kind):
    """per-day nll and d nll / d eta vs the oracle's scipy-derived formulas; Fisher weights are the
    expectations of the oracle's second derivatives (checked by Monte Carlo for one setting)."""
    rng = np.random.default_rng(2)
    n = 2000
    ea = rng.normal(0, 0.4, n)
    eb = rng.normal(0.5, 0.4, n)
    if kind == "normal":
        y = rng.normal(0.3, 1.0, n)
    elif kind == "beta":
        y = rng.uniform(0.02, 0.98, n)
    elif kind == "bernoulli":
        y = (rng.uniform(size=n) < 0.4).astype(float)
    else:
        y = rng.gamma(2.0, 0.6, n) + 1e-3
    got = hm.family(kind, y, ea, eb)
    nll, ra, rb, waa, wab, wbb = oracle.family_terms(kind, y, ea, eb if kind != "bernoulli" else None)
    np.testing.assert_allclose(got[:, 0], nll, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got[:, 1], ra, rtol=1e-11, atol=1e-12)
    if rb is not None:
        np.testing.assert_allclose(got[:, 2], rb, rtol=1e-10, atol=1e-11)
    got32 = hm.family(kind, y, ea, eb, f32=True)
    np.testing.assert_allclose(got32[:, 0], nll, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(got32[:, 1], ra, rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("kind", ["normal", "gamma", "beta", "weibull"])
def test_fisher_weights_are_expected_hessians(kind):
    rng = np.random.default_rng(3)
    ea, eb, n = 0.3, 0.7, 400000
    if kind == "normal":
        y = rng.normal(ea, np.exp(eb), n)
    elif kind == "gamma":
        a = np.exp(2 * eb)
        y = rng.gamma(a, np.exp(ea) / a, n)
    elif kind == "weibull":
        y = np.exp(ea) * rng.weibull(np.exp(eb), n)
    else:
        mu, phi = 1 / (1 + np.exp(-ea)), np.exp(eb)
        y = np.clip(rng.beta(mu * phi, (1 - mu) * phi, n), 1e-12, 1 - 1e-12)
    _, _, _, waa, wab, wbb = oracle.family_terms(kind, y, np.full(n, ea), np.full(n, eb))
    w = hm.family(kind, y[:1], [ea], [eb])[0]
    for got, mc in ((w[3], waa), (w[4], wab), (w[5], wbb)):
        assert abs(got - mc.mean()) <= 6 * mc.std() / np.sqrt(n) + 1e-9


def _to_ref_layout(cell, thetas):
    ci = hm.canon_index(cell.family, helpers.MODES)
    out = np.zeros(len(ci))
    for j, (s, c) in enumerate(ci):
        out[j] = thetas[s][c]
    return out


def _terms(cell):
    return ["bernoulli", "gamma"] if cell.family == "bernoulli_gamma" else [cell.family]


@pytest.mark.parametrize("name", helpers.ONE_PER_VARIABLE)
def test_pass_and_assemble_against_oracle(name):
    """NLL, gradient and Fisher information through pass_impl + step_impl (moment form, rotated frames)."""
    c = helpers.Cell(name)
    th = c.d["theta_probe"]
    f, g, _ = c.problem.nll(th, 2)
    ci = hm.canon_index(c.family, helpers.MODES)
    tot, gg = 0.0, np.zeros(c.problem.n)
    for slot, term in enumerate(_terms(c)):
        thc = np.zeros(hm.NP_CANON)
        for j, (s, cn) in enumerate(ci):
            if s == slot:
                thc[cn] = th[j]
        nll, grad, fisher = hm.nll_grad_cell(term, helpers.MODES, c.ys, c.days, c.gmt, 0, c.n_fit, thc, fp64=True)
        nll32, grad32, _ = hm.nll_grad_cell(term, helpers.MODES, c.ys, c.days, c.gmt, 0, c.n_fit, thc, fp64=False)
        assert abs(nll32 - nll) <= 2e-7 * abs(nll)
        assert np.abs(grad32 - grad).max() <= 2e-5 * np.abs(grad).max() + 1e-2
        assert np.all(np.linalg.eigvalsh(fisher) > 0)
        tot += nll
        for j, (s, cn) in enumerate(ci):
            if s == slot:
                gg[j] = grad[cn]
    np.testing.assert_allclose(tot, f, rtol=1e-12)
    np.testing.assert_allclose(tot, c.d["ref_f_probe"], rtol=1e-6 if c.variable == "pr" else 1e-12)
    np.testing.assert_allclose(gg, g, rtol=1e-9, atol=1e-7)


@pytest.mark.parametrize("name", helpers.FIXTURES)
@pytest.mark.parametrize("fp64", [False, True])
def test_fit_against_oracle(name, fp64):
    """The Fisher-scoring LM iteration reaches the oracle's exact MAP; logp no worse than the reference."""
    c = helpers.Cell(name)
    ref = oracle.fit_newton(c.problem)
    thetas, logp, npass = [], 0.0, 0
    for term in _terms(c):
        r = hm.fit_cell(term, helpers.MODES, c.ys, c.days, c.gmt, 0, c.n_fit, fp64=fp64)
        assert r["status"] == 0
        thetas.append(r["theta"])
        logp += r["logp"]
        npass += r["n_pass"]
    theta = _to_ref_layout(c, thetas)
    assert npass <= 25
    np.testing.assert_allclose(logp, ref["logp"], rtol=1e-9)
    assert np.abs(theta - ref["params"]).max() < 2e-5
    ref_nll = -float(c.d["ref_logp"])
    assert -logp <= ref_nll + 1e-6 * abs(ref_nll)


def test_fit_zero_start_and_shifted_phase_origin():
    """zero start like the reference (model_scipy.py:113, 234) and a cell whose first valid day is not
    day 0 (phase-origin quirk)."""
    c = helpers.Cell("tas_lat50.75_lon9.25")
    ys = c.ys.copy()
    ys[:37] = np.nan
    prob = oracle.build_problem("normal", ys, c.gmt, c.days, helpers.MODES, slice(0, c.n_fit))
    ref = oracle.fit_newton(prob)
    for zero in (True, False):
        r = hm.fit_cell("normal", helpers.MODES, ys, c.days, c.gmt, 0, c.n_fit, zero_init=zero)
        assert r["status"] == 0
        np.testing.assert_allclose(r["logp"], ref["logp"], rtol=1e-9)
        assert np.abs(_to_ref_layout(c, [r["theta"]]) - ref["params"]).max() < 2e-5


@pytest.mark.parametrize("modes", [1, 2, 3])
def test_fit_modes_sweep(modes):
    c = helpers.Cell("sfcWind_lat50.75_lon9.25")
    prob = oracle.build_problem("weibull", c.ys, c.gmt, c.days, modes, slice(0, c.n_fit))
    ref = oracle.fit_newton(prob)
    r = hm.fit_cell("weibull", modes, c.ys, c.days, c.gmt, 0, c.n_fit)
    ci = hm.canon_index("weibull", modes)
    theta = np.array([r["theta"][cn] for _, cn in ci])
    assert r["status"] == 0
    np.testing.assert_allclose(r["logp"], ref["logp"], rtol=1e-9)
    assert np.abs(theta - ref["params"]).max() < 2e-5


@pytest.mark.parametrize("name", helpers.ONE_PER_VARIABLE)
def test_quantile_map_against_reference(name):
    c = helpers.Cell(name)
    d = c.d
    r = hm.qmap_cell(c.family, c.kind, c.post, helpers.MODES, d["y"], c.ys, c.days, c.gmt, d["ref_params"],
                     c.scaling.get("datamin", 0.0), c.scaling.get("scale", 1.0), c.uniforms)
    idx = d["sample_idx"]
    np.testing.assert_allclose(r["cfact"][idx], d["ref_cfact"], rtol=1e-6 if c.variable == "pr" else 1e-11, equal_nan=True)
    assert np.array_equal(r["replaced"][idx], helpers.replaced_codes(d["ref_replaced"]))
    for i, k in enumerate(oracle.FAMILY_PARAMS[c.family]):
        np.testing.assert_allclose(r["par_ref"][idx, i], d[f"ref_{k[0]}"], rtol=1e-12)
        np.testing.assert_allclose(r["par_cf"][idx, i], d[f"ref_{k[0]}_cfact"], rtol=1e-12)
    if c.variable == "pr":
        assert np.array_equal(r["cfact"] == 0, d["ref_cfact"] == 0)


def test_quantile_map_tails_and_replacement():
    """+-Inf / NaN from the map fall back to the original value with the reference's flags (detrend.py:397-411)."""
    c = helpers.Cell("tas_lat50.75_lon9.25")
    ys = c.ys[:64].copy()
    y = c.d["y"][:64].copy()
    ys[3], ys[5], ys[7] = 40.0, -40.0, np.nan  # cdf -> 1, 0, NaN
    r = hm.qmap_cell("normal", "unity", None, helpers.MODES, y, ys, c.days[:64], c.gmt[:64], c.d["ref_params"],
                     c.scaling["datamin"], c.scaling["scale"])
    assert list(r["replaced"][[3, 5, 7]]) == [2, 3, 1]
    assert r["cfact"][3] == np.float64(y[3]) and r["cfact"][5] == np.float64(y[5]) and r["cfact"][7] == np.float64(y[7])
    ref = oracle.estimate_params("normal", c.d["ref_params"], c.gmt[:64], c.days[:64], helpers.MODES)
    cf = oracle.estimate_params("normal", c.d["ref_params"], 0 * c.gmt[:64], c.days[:64], helpers.MODES)
    cs = oracle.quantile_map("tas", ys, ref, cf)
    cfact, rep = oracle.replace_invalid(oracle.rescale("tas", cs, c.scaling), y)
    np.testing.assert_allclose(r["cfact"], cfact, rtol=1e-12, equal_nan=True)
    assert np.array_equal(r["replaced"], helpers.replaced_codes(rep))


def test_mt19937_matches_numpy_legacy():
    for seed in (0, 1, 12345, 2**32 - 1, oracle.cell_seed(0, 50.75, 9.25)):
        np.random.seed(seed)
        assert np.array_equal(hm.mt19937(seed, 1500), np.random.rand(1500))
