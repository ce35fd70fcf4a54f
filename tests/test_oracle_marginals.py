"""Checks of oracle/marginals_oracle.py (the WLS marginals restatement).  OpenMx is not available
(parity unpinned), so every building block is pinned independently: the bivariate normal CDF
against scipy, stage 1 against a generic optimiser, every analytic score against finite
differences of an independently written log-likelihood, and Gamma against Monte Carlo."""
import numpy as np
import pandas as pd
import pytest
from scipy import optimize, stats
from scipy.special import ndtr

import marginals_oracle as mo
from gwsem_b200.gwas import model_columns
from gwsem_b200.model import buildOneFac, flatten


def test_bivariate_normal_cdf():
    rng = np.random.default_rng(0)
    for rho in (-0.8, -0.3, 0.0, 0.2, 0.6, 0.9):
        mvn = stats.multivariate_normal([0, 0], [[1, rho], [rho, 1]])
        for _ in range(5):
            h, k = rng.normal(size=2) * 1.5
            assert mo.bvn_cdf(h, k, rho) == pytest.approx(mvn.cdf([h, k]), abs=2e-8)


def sim(n, seed, K=2):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, K))
    f = rng.normal(size=n)
    y1 = 0.8 * f + X @ [0.3, -0.2][:K] + rng.normal(size=n)
    y2 = 0.6 * f + X @ [-0.1, 0.25][:K] + rng.normal(size=n)
    c = 0.7 * f + X @ [0.2, 0.1][:K] + rng.normal(size=n) + 0.5
    return X, np.searchsorted([-.6, .4], y1), np.searchsorted([-.9, 0, .7], y2), c


def test_probit_stage1_is_the_mle():
    X, o1, _, _ = sim(1500, 1)
    v = mo.Ord(o1, X, 3)
    assert np.abs(v.sc.sum(0)).max() < 1e-8                      # score vanishes

    def nll(par):
        th = np.concatenate([[-12], par[:2], [12]])
        eta = X @ par[2:]
        return -np.log(ndtr(th[o1 + 1] - eta) - ndtr(th[o1] - eta)).sum()

    alt = optimize.minimize(nll, np.r_[-.5, .5, 0, 0], method="BFGS", options=dict(gtol=1e-8))
    assert np.allclose(np.r_[v.th, v.gamma], alt.x, atol=2e-5)


def loglik_pair(kind, a_par, b_par, rho, data):
    """Independent per-person bivariate log-likelihoods (used only for finite differences)."""
    X = data["X"]
    if kind == "polyserial":
        y, codes = data["c"], data["o"]
        K = X.shape[1]
        beta, var = a_par[:K + 1], a_par[K + 1]
        u = (y - beta[0] - X @ beta[1:]) / np.sqrt(var)
        C = len(b_par) - K + 1
        th = np.concatenate([[-12], b_par[:C - 1], [12]])
        eta = X @ b_par[C - 1:]
        s = np.sqrt(1 - rho ** 2)
        a1, a0 = (th[codes + 1] - eta - rho * u) / s, (th[codes] - eta - rho * u) / s
        return -0.5 * np.log(var) - 0.5 * u ** 2 + np.log(ndtr(a1) - ndtr(a0))
    if kind == "polychoric":
        ca, cb = data["oa"], data["ob"]
        K = X.shape[1]
        Ca, Cb = len(a_par) - K + 1, len(b_par) - K + 1
        tha = np.concatenate([[-12], a_par[:Ca - 1], [12]])
        thb = np.concatenate([[-12], b_par[:Cb - 1], [12]])
        ea, eb = X @ a_par[Ca - 1:], X @ b_par[Cb - 1:]
        P = (mo.bvn_cdf(tha[ca + 1] - ea, thb[cb + 1] - eb, rho) - mo.bvn_cdf(tha[ca] - ea, thb[cb + 1] - eb, rho)
             - mo.bvn_cdf(tha[ca + 1] - ea, thb[cb] - eb, rho) + mo.bvn_cdf(tha[ca] - ea, thb[cb] - eb, rho))
        return np.log(P)
    ya, yb = data["a"], data["b"]
    K = X.shape[1]
    u = (ya - a_par[0] - X @ a_par[1:K + 1]) / np.sqrt(a_par[K + 1])
    v = (yb - b_par[0] - X @ b_par[1:K + 1]) / np.sqrt(b_par[K + 1])
    s2 = 1 - rho ** 2
    return -0.5 * np.log(a_par[K + 1] * b_par[K + 1]) - 0.5 * np.log(s2) - (u * u - 2 * rho * u * v + v * v) / (2 * s2)


def fd(f, x, k, h=1e-6):
    e = np.zeros(len(x))
    e[k] = h
    return (f(x + e) - f(x - e)) / (2 * h)


@pytest.mark.parametrize("kind", ["pearson", "polyserial", "polychoric"])
def test_pair_scores_match_finite_differences(kind):
    X, o1, o2, c = sim(400, 2)
    c2 = c * 0.5 + np.random.default_rng(3).normal(size=400)
    if kind == "pearson":
        a, b = mo.Cont(c, X), mo.Cont(c2, X)
        rho, sc, da, db = mo.pair_pearson(a, b)
        data = dict(X=X, a=c, b=c2)
        pa, pb = np.r_[a.beta, a.var], np.r_[b.beta, b.var]
    elif kind == "polyserial":
        a, b = mo.Cont(c, X), mo.Ord(o1, X, 3)
        rho, sc, da, db = mo.pair_polyserial(a, b)
        data = dict(X=X, c=c, o=o1)
        pa, pb = np.r_[a.beta, a.var], np.r_[b.th, b.gamma]
    else:
        a, b = mo.Ord(o1, X, 3), mo.Ord(o2, X, 4)
        rho, sc, da, db = mo.pair_polychoric(a, b)
        data = dict(X=X, oa=o1, ob=o2)
        pa, pb = np.r_[a.th, a.gamma], np.r_[b.th, b.gamma]
    assert abs(sc.sum()) < 1e-6 * len(sc)                        # rho solves its score equation
    num = (loglik_pair(kind, pa, pb, rho + 1e-6, data) - loglik_pair(kind, pa, pb, rho - 1e-6, data)) / 2e-6
    assert np.allclose(sc, num, atol=2e-6 * (1 + np.abs(num).max()))
    for k in range(len(pa)):
        num = fd(lambda p: loglik_pair(kind, p, pb, rho, data), pa, k)
        assert np.allclose(da[:, k], num, atol=5e-6 * (1 + np.abs(num).max())), (kind, "a", k)
    for k in range(len(pb)):
        num = fd(lambda p: loglik_pair(kind, pa, p, rho, data), pb, k)
        assert np.allclose(db[:, k], num, atol=5e-6 * (1 + np.abs(num).max())), (kind, "b", k)


def test_gamma_is_calibrated_by_monte_carlo():
    reps, n = 150, 600
    S, G = [], []
    for r in range(reps):
        X, o1, o2, c = sim(n, 100 + r, K=1)
        # a normal stand-in for the genotype column: with outer-product A11 (the lavaan / OpenMx
        # convention restated here) the variance statistic's own acov is only calibrated under normality
        g = np.random.default_rng(500 + r).normal(0.7, 0.65, n)
        st = mo.marginals_stats([g, o1.astype(float), c], ["cont", "ord", "cont"], [0, 3, 0], X,
                                np.array([[False], [True], [True]]))
        S.append(st["s"])
        G.append(st["nacov"] / n)
    emp = np.cov(np.array(S).T)
    avg = np.mean(G, 0)
    ratio = np.diag(avg) / np.diag(emp)
    assert np.all((ratio > 0.7) & (ratio < 1.4)), ratio


def test_fit_recovers_c3_like_parameters():
    rng = np.random.default_rng(4)
    n, K = 3000, 2
    g = rng.binomial(2, .3, n).astype(float)
    X = rng.normal(size=(n, K))
    F = 0.25 * g + rng.normal(size=n)
    ph = pd.DataFrame({f"pc{k + 1}": X[:, k] for k in range(K)})
    for j, (lam, cuts) in enumerate(zip([.8, .7, .6], [[-.5, .5], [-.8, 0, .8], [-1, -.3, .3, 1.]])):
        ystar = lam * F + X @ (rng.normal(size=K) * 0.2) + rng.normal(size=n)
        ph[f"i{j + 1}"] = pd.Categorical(np.searchsorted(cuts, ystar), ordered=True)
    model = buildOneFac(ph, ["i1", "i2", "i3"], covariates=["pc1", "pc2"])
    flat = flatten(model)
    data, ncat = model_columns(model)
    r = mo.fit_snp_marginals(flat, data, g, ncat)
    assert r["status"] == "OK"
    k = flat["par_names"].index("snp_to_F")
    assert abs(r["est"][k] - 0.25) < 3.5 * np.sqrt(r["vcov"][k, k])
    lam_hat = [r["est"][flat["par_names"].index(f"lambda_i{j}")] for j in (1, 2, 3)]
    assert np.allclose(lam_hat, [.8, .7, .6], atol=0.15)
