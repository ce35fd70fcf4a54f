"""The compiled restatement of the WLS marginals fit (oracle/marginals_port.c, the CPU baseline of bench.py) against the
numpy oracle it restates (oracle/marginals_oracle.py, itself pinned by the reference's known answers): same estimates,
standard errors, status and catch on ordinal, continuous and mixed models, with covariates, with missing calls and on
a monomorphic SNP."""
import numpy as np
import pandas as pd
import pytest

import marginals_oracle as mo
from gwsem_b200 import synth
from gwsem_b200.gwas import model_columns
from gwsem_b200.model import buildItem, buildOneFac, buildOneFacRes, flatten
from port_lib import MarginalsPort


def cohort(n, seed, n_items, n_cov, ordinal):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, n_cov))
    f = rng.normal(size=n)
    ph = pd.DataFrame({f"pc{k + 1}": X[:, k] for k in range(n_cov)})
    cuts = [[-.5, .5], [-.8, 0, .8], [-1, -.3, .3, 1.], [0.], [-.4, .6]]
    for j in range(n_items):
        ystar = (0.8 - 0.1 * j) * f + X @ (rng.normal(size=n_cov) * 0.2) + rng.normal(size=n)
        ph[f"i{j + 1}"] = pd.Categorical(np.searchsorted(cuts[j % len(cuts)], ystar), ordered=True) if ordinal[j] else ystar + 0.3
    return ph


CASES = {
    "onefac_ordinal_pcs": dict(n_items=3, n_cov=3, ordinal=[True, True, True], build="onefac"),
    "onefac_mixed": dict(n_items=4, n_cov=2, ordinal=[True, False, True, False], build="onefac"),
    "onefacres_continuous": dict(n_items=4, n_cov=2, ordinal=[False] * 4, build="onefacres"),
    "item_ordinal_exogenous_snp": dict(n_items=1, n_cov=2, ordinal=[True], build="item"),
}


@pytest.mark.parametrize("case", list(CASES))
def test_port_matches_numpy_oracle(case):
    cfg = CASES[case]
    n = 1200
    ph = cohort(n, 31, cfg["n_items"], cfg["n_cov"], cfg["ordinal"])
    items = [f"i{j + 1}" for j in range(cfg["n_items"])]
    covs = [f"pc{k + 1}" for k in range(cfg["n_cov"])]
    if cfg["build"] == "onefac":
        model = buildOneFac(ph, items, covariates=covs)
    elif cfg["build"] == "onefacres":
        model = buildOneFacRes(ph, items, covariates=covs)
    else:
        model = buildItem(ph, items[0], covariates=covs)
    flat = flatten(model)
    assert flat["continuous_type"] == "marginals"
    data, ncat = model_columns(model)
    codes = synth.simulate_hardcalls(n, 4, seed=7, maf_range=(0.15, 0.5), missing_rate=0.0)
    codes[1] = 0                                     # monomorphic: minVariance
    codes[2, ::17] = 3                               # missing calls: naAction = 'omit'
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    port = MarginalsPort(flat, data, ncat)
    for g in geno:
        want = mo.fit_snp_marginals(flat, data, g, ncat)
        got = port.fit(g)
        assert got["catch1"] == want["catch1"] and got["status"] == want["status"] and got["n"] == want["n"]
        if want["catch1"]:
            continue
        se = np.sqrt(np.diag(want["vcov"]))
        assert np.max(np.abs(got["est"] - want["est"]) / np.maximum(np.abs(want["est"]), se)) < 1e-6
        assert np.allclose(np.sqrt(np.diag(got["vcov"])), se, rtol=1e-6, atol=0)
        assert abs(got["fit"] - want["fit"]) < 1e-6 * (1 + abs(want["fit"]))
