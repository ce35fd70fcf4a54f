"""Checks of the numpy fit restatement (oracle/fit_oracle.py).  The reference delegates this
arithmetic to OpenMx, which is not available here, so the oracle is pinned by construction:
closed forms, finite differences, an independent optimiser and simulation recovery."""
import numpy as np
import pandas as pd
import pytest
from scipy import optimize

import fit_oracle as fo
from gwsem_b200.model import buildItem, buildOneFac, buildOneFacRes, flatten


def cohort(n, seed):
    rng = np.random.default_rng(seed)
    g = rng.binomial(2, .3, n).astype(float)
    f = rng.normal(size=n) + 0.1 * g
    ph = pd.DataFrame({f"i{k}": (0.5 + 0.1 * k) * f + rng.normal(size=n) * (1 + 0.1 * k) + 0.05 * k * g
                       for k in range(1, 6)})
    return ph, g


def test_saturated_item_closed_form():
    ph, g = cohort(4000, 3)
    flat = flatten(buildItem(ph, "i1"))
    r = fo.fit_snp_cumulants(flat, {"i1": ph["i1"].to_numpy()}, g)
    c = np.cov(np.stack([ph["i1"], g]))
    want = [c[0, 1] / c[1, 1], c[0, 0] - c[0, 1] ** 2 / c[1, 1], c[1, 1]]
    assert r["status"] == "OK" and np.allclose(r["est"], want, rtol=1e-9)
    # ADF standard error of the slope equals the heteroskedasticity-robust (HC0) OLS one
    d = np.stack([ph["i1"], g], 1) - [ph["i1"].mean(), g.mean()]
    e = d[:, 0] - want[0] * d[:, 1]
    hc0 = np.sqrt((d[:, 1] ** 2 * e ** 2).sum()) / (d[:, 1] ** 2).sum()
    assert np.sqrt(r["vcov"][0, 0]) == pytest.approx(hc0, rel=2e-3)


def test_analytic_jacobian_matches_finite_differences():
    ph, g = cohort(500, 4)
    items = list(ph.columns)
    for model in (buildOneFac(ph, items), buildOneFacRes(ph, items, factor=False)):
        ram = fo.FlatRAM(flatten(model))
        th = np.array(flatten(model)["par_start"]) + np.linspace(-.2, .3, ram.P)
        J = fo.jac_cumulants(ram, th)
        for k in range(ram.P):
            h = np.zeros(ram.P)
            h[k] = 1e-6
            fd = (fo.sigma_cumulants(ram, th + h) - fo.sigma_cumulants(ram, th - h)) / 2e-6
            assert np.allclose(J[:, k], fd, atol=1e-7)


def test_gauss_newton_agrees_with_generic_optimiser():
    ph, g = cohort(3000, 5)
    items = list(ph.columns)
    flat = flatten(buildOneFacRes(ph, items))
    cols = {c: ph[c].to_numpy() for c in items}
    r = fo.fit_snp_cumulants(flat, cols, g)
    ram = fo.FlatRAM(flat)
    s, gamma, n = fo.cumulants_stats(fo.snp_columns(flat, cols, g))
    W = n * np.linalg.inv(gamma)
    obj = lambda t: (s - fo.sigma_cumulants(ram, t)) @ W @ (s - fo.sigma_cumulants(ram, t))
    alt = optimize.minimize(obj, flat["par_start"], method="L-BFGS-B", options=dict(maxiter=5000, ftol=1e-15, gtol=1e-10))
    assert r["status"] == "OK" and r["fit"] <= alt.fun + 1e-8
    assert np.allclose(r["est"], alt.x, atol=2e-4)


def test_simulation_recovery_and_se_calibration():
    n, reps = 2000, 60
    z = []
    for rep in range(reps):
        rng = np.random.default_rng(100 + rep)
        g = rng.binomial(2, .3, n).astype(float)
        f = rng.normal(size=n) + 0.15 * g
        ph = pd.DataFrame({f"i{k}": 0.7 * f + rng.normal(size=n) for k in range(1, 5)})
        flat = flatten(buildOneFac(ph, list(ph.columns)))
        r = fo.fit_snp_cumulants(flat, {c: ph[c].to_numpy() for c in ph.columns}, g)
        k = flat["par_names"].index("snp_to_F")
        sign = np.sign(r["est"][flat["par_names"].index("lambda_i1")])
        z.append((sign * r["est"][k] - 0.15) / np.sqrt(r["vcov"][k, k]))
    z = np.array(z)
    assert abs(z.mean()) < 0.45 and 0.7 < z.std() < 1.35


def test_min_variance_and_na_action():
    ph, g = cohort(300, 6)
    flat = flatten(buildItem(ph, "i1"))
    g0 = np.zeros(300)
    g0[:2] = 1                                   # variance below 2*0.01*0.99
    r = fo.fit_snp_cumulants(flat, {"i1": ph["i1"].to_numpy()}, g0)
    assert "observed variance less" in r["catch1"]   # tests/testthat/test-covariate.R:62
    g2 = g.copy()
    g2[::7] = np.nan
    r = fo.fit_snp_cumulants(flat, {"i1": ph["i1"].to_numpy()}, g2)
    assert r["n"] == int(np.isfinite(g2).sum()) and r["status"] == "OK"
