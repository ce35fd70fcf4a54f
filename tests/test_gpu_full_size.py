"""Size-independent properties at the cohort sizes BASELINE.json quotes (C2: 100k people, C3: 50k people), where
the numpy oracles only afford a couple of SNPs:
  * C2 (buildItem, one continuous phenotype, WLS cumulants) is a saturated model with a closed form
    (SURVEY.md section 8a row O3): snp_to_y = S_ys / S_ss, snp_res = S_ss, y_res = S_yy - S_ys^2 / S_ss;
  * a row never depends on its neighbours (the reference restarts every SNP from the original starts,
    R/model.R:188): any sub-batch, in any order, gives the same bits;
  * PLINK 1 coding of the same genotypes gives the same bits;
  * C3 (buildOneFac, 3 ordinal items + 5 PCs, WLS marginals): two SNPs against oracle/marginals_oracle.py within the
    north-star tolerances, everything else through the two invariances."""
import numpy as np
import pandas as pd
import pytest

from gwsem_b200 import synth
from gwsem_b200.gwas import model_columns
from gwsem_b200.model import buildItem, buildOneFac, flatten

pytestmark = pytest.mark.gpu

EST_RTOL, SE_RTOL = 1e-4, 1e-3   # BASELINE.json north_star tolerances


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def test_c2_closed_form_and_row_independence(engine):
    from gwsem_b200.engine import Model
    n, m = 100_000, 48
    rng = np.random.default_rng(2001)
    codes = synth.simulate_hardcalls(n, m, seed=1001, missing_rate=0.0)
    codes[7] = synth.simulate_hardcalls(n, 1, seed=5, missing_rate=0.005)[0]   # one SNP with missing calls
    y = 0.05 * np.where(codes[3] == 3, 0, codes[3]) + rng.normal(size=n)
    model = buildItem(pd.DataFrame({"y": y}), "y")
    flat = flatten(model)
    names = flat["par_names"]
    gm = Model(engine, flat, {"y": y})
    res = gm.fit_2bit(synth.pack_2bit(codes))
    assert (res.status == 0).all() and (res.catch_code == 0).all()
    g = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    for i in range(m):
        ok = ~np.isnan(g[i])
        gi, yi = g[i][ok], y[ok]
        assert res.n_used[i] == ok.sum()
        c = np.cov(np.vstack([yi, gi]))          # N - 1 denominators (conventions)
        want = {"snp_to_y": c[0, 1] / c[1, 1], "snp_res": c[1, 1], "y_res": c[0, 0] - c[0, 1] ** 2 / c[1, 1]}
        for k, nm in enumerate(names):
            if nm in want:
                assert abs(res.est[i, k] - want[nm]) <= 1e-6 * max(abs(want[nm]), 1e-3), (i, nm, res.est[i, k], want[nm])
    # any sub-batch in any order: the same bits, row for row
    pick = np.array([40, 3, 7, 0, 47, 21])
    sub = gm.fit_2bit(synth.pack_2bit(codes[pick]))
    assert same_bits(sub.est, res.est[pick]) and same_bits(sub.vcov, res.vcov[pick]) and same_bits(sub.maf, res.maf[pick])
    bed = gm.fit_2bit(synth.pack_2bit(np.array([3, 2, 0, 1], np.uint8)[codes]), plink1=True)
    assert same_bits(bed.est, res.est) and same_bits(bed.vcov, res.vcov)
    gm.close()


def c3_cohort(n, seed=2002):
    rng = np.random.default_rng(seed)
    K = 5
    X = rng.normal(size=(n, K))
    F = rng.normal(size=n) + X @ np.full(K, 0.1)
    ph = pd.DataFrame({f"pc{k + 1}": X[:, k] for k in range(K)})
    cuts = [[-.43, .43], [-.67, 0, .67], [-.84, -.25, .25, .84]]
    for j, lam in enumerate([.8, .7, .6]):
        ystar = lam * F + X @ (rng.normal(size=K) * 0.1) + rng.normal(size=n)
        ph[f"i{j + 1}"] = pd.Categorical(np.searchsorted(cuts[j], ystar), ordered=True)
    return ph


def test_c3_oracle_spot_check_and_row_independence(engine):
    import marginals_oracle as mo
    from gwsem_b200.engine import STATUS_TEXT, Model
    n, m = 50_000, 24
    ph = c3_cohort(n)
    model = buildOneFac(ph, ["i1", "i2", "i3"], covariates=[f"pc{k + 1}" for k in range(5)])
    flat = flatten(model)
    data, ncat = model_columns(model)
    codes = synth.simulate_hardcalls(n, m, seed=1002, missing_rate=0.0)
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    assert (res.status == 0).all() and (res.catch_code == 0).all()
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    for i in (0, 13):
        w = mo.fit_snp_marginals(flat, data, geno[i], ncat)
        assert STATUS_TEXT[int(res.status[i])] == w["status"] and res.n_used[i] == w["n"]
        se_w, se_g = np.sqrt(np.diag(w["vcov"])), np.sqrt(np.diag(res.vcov[i]))
        assert np.all(np.abs(res.est[i] - w["est"]) <= EST_RTOL * np.maximum(np.abs(w["est"]), se_w)), i
        assert np.allclose(se_g, se_w, rtol=SE_RTOL, atol=0), i
    pick = np.array([23, 0, 13, 5])
    sub = gm.fit_2bit(synth.pack_2bit(codes[pick]))
    assert same_bits(sub.est, res.est[pick]) and same_bits(sub.vcov, res.vcov[pick])
    bed = gm.fit_2bit(synth.pack_2bit(np.array([3, 2, 0, 1], np.uint8)[codes]), plink1=True)
    assert same_bits(bed.est, res.est) and same_bits(bed.vcov, res.vcov)
    gm.close()
