"""CPU checks of oracle/ml_oracle.py (the ML / FIML restatement):
the augmented-moment form equals the row-by-row definition of FIML, and buildItem under ML is
the ML regression of the item on the genotype (closed form)."""
import numpy as np
import pandas as pd

import ml_oracle as mo
from gwsem_b200.gwas import model_columns
from gwsem_b200.model import buildItem, buildOneFac, flatten


def test_moment_form_equals_rowwise_fiml():
    rng = np.random.default_rng(0)
    n = 300
    F = rng.normal(size=n)
    ph = pd.DataFrame({"pc1": rng.normal(size=n), "E": rng.integers(0, 2, n).astype(float)})
    for k, lam in enumerate([.8, .7, .6, .9], 1):
        ph[f"i{k}"] = lam * F + 0.2 * ph["pc1"] + rng.normal(size=n) + 0.5
    model = buildOneFac(ph, ["i1", "i2", "i3", "i4"], gxe="E", covariates=["pc1"], fitfun="ML")
    flat = flatten(model)
    data, _ = model_columns(model)
    g = rng.binomial(2, 0.3, n).astype(float)
    g[:9] = np.nan
    w = mo.fit_snp_ml(flat, data, g)
    assert w["status"] == "OK" and w["n"] == n - 9
    assert abs(w["fit"] - mo.neg2ll_rows(flat, w["est"], data, g)) < 1e-9 * abs(w["fit"])
    assert np.all(np.diag(w["vcov"]) > 0)
    # the gradient vanishes at the estimate and the analytic gradient matches differences of the objective
    ram = mo.FlatRAM(flat)
    Y = np.column_stack([g if m == "snp" else g * data["E"] if m == "snp_E" else data[m] for m in flat["manifest"]])
    W = np.column_stack([np.ones(n), data["pc1"]])
    ok = np.isfinite(g)
    dep = np.array([m.startswith("snp") for m in flat["manifest"]])
    groups = [mo.Group(np.arange(Y.shape[1]), Y[ok], W[ok]), mo.Group(np.nonzero(~dep)[0], Y[~ok][:, ~dep], W[~ok])]
    theta = w["est"] * 1.01 + 0.01
    f0, g0 = mo.neg2ll(ram, flat, theta, groups, grad=True)
    for k in (0, 3, 10, len(theta) - 1):
        h = 1e-6
        e = np.zeros(len(theta)); e[k] = h
        num = (mo.neg2ll(ram, flat, theta + e, groups) - mo.neg2ll(ram, flat, theta - e, groups)) / (2 * h)
        assert abs(num - g0[k]) <= 1e-5 * max(1.0, abs(num))


def test_item_under_ml_is_regression():
    rng = np.random.default_rng(1)
    n = 800
    g = rng.binomial(2, 0.35, n).astype(float)
    g[:11] = np.nan
    ph = pd.DataFrame({"y": 0.2 * np.nan_to_num(g) + rng.normal(size=n) + 1.5})
    model = buildItem(ph, "y", fitfun="ML")
    flat = flatten(model)
    data, _ = model_columns(model)
    w = mo.fit_snp_ml(flat, data, g)
    ok = np.isfinite(g)
    X = np.column_stack([np.ones(ok.sum()), g[ok]])
    beta = np.linalg.lstsq(X, ph["y"][ok], rcond=None)[0]
    r = ph["y"][ok] - X @ beta
    s2 = r @ r / ok.sum()
    names = flat["par_names"]
    assert abs(w["est"][names.index("snp_to_y")] - beta[1]) < 1e-7
    assert abs(w["est"][names.index("y_res")] - s2) < 1e-7
    k = names.index("snp_to_y")
    assert abs(w["vcov"][k, k] / (s2 * np.linalg.inv(X.T @ X)[1, 1]) - 1) < 1e-5
    mono = mo.fit_snp_ml(flat, data, np.zeros(n))
    assert "observed variance less" in mono["catch1"]
