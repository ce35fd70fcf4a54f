"""Host-side mirror of the R API (reference R/model.R): structure checks taken from the
reference's own tests (tests/testthat/test-model.R, test-covariate.R, test-gxe.R)."""
import numpy as np
import pandas as pd
import pytest

from gwsem_b200.model import buildItem, buildOneFac, buildOneFacRes, buildTwoFac, flatten


@pytest.fixture()
def pheno():
    rng = np.random.default_rng(1)
    ph = pd.DataFrame({f"i{k}": rng.normal(size=60) for k in range(1, 8)})
    ph["i1"] = pd.Categorical(pd.cut(ph["i1"], [-np.inf, -.4, .4, np.inf]), ordered=True)
    ph["i2"] = pd.Categorical(pd.cut(ph["i2"], [-np.inf, 0, np.inf]), ordered=True)
    ph["covar1"] = rng.normal(size=60)
    ph["covar2"] = rng.normal(size=60)
    return ph


def test_dots_barrier(pheno):
    # test-model.R:34-47: "Rejected are any values passed in the '...' argument"
    for fn, args in ((buildItem, ("i3", None)), (buildOneFac, (["i3", "i4"], None)),
                     (buildOneFacRes, (["i3", "i4"], False, None, None)),
                     (buildTwoFac, (["i3", "i4"], ["i5", "i6"], None))):
        with pytest.raises(ValueError, match="Rejected are any values"):
            fn(pheno, *args, "oops")


def test_build_item_continuous_is_saturated_cumulants(pheno):
    m = buildItem(pheno, "i3")
    assert m.manifestVars == ["i3", "snp"] and m.continuousType == "cumulants" and not m.hasMeans
    names, start, lb, _ = m.parameters()
    assert names == ["snp_to_i3", "i3_res", "snp_res"]
    assert list(start) == [0, 1, 1] and np.isnan(lb[0]) and list(lb[1:]) == [.001, .001]
    assert m.minVariance == pytest.approx(2 * .01 * .99)


def test_build_item_ordinal_is_exogenous_marginals(pheno):
    m = buildItem(pheno, "i1")
    assert m.manifestVars == ["i1"] and m.latentVars == ["snp"] and m.continuousType == "marginals"
    assert list(m.thresholds) == ["i1"] and m.thresholds["i1"]["labels"] == ["i1_Thr_1", "i1_Thr_2"]


def test_one_fac_with_exogenous_covariates(pheno):
    items = [f"i{k}" for k in range(1, 8)]
    m = buildOneFac(pheno, items, covariates=["covar1", "covar2"], exogenous=True)
    assert m.M["labels"][0, m.vars.index("covar1")] == "data.covar1"     # test-covariate.R:37
    names = m.parameters()[0]
    for c in ("covar1", "covar2"):
        for i in items:
            assert f"{c}_to_{i}" in names
    assert m.exoFree.loc["snp"].sum() == 0 and m.exoFree.loc[items].all().all()
    assert "i1_res" not in names and "i3_res" in names and "i1Mean" not in names   # ordinal: fixed
    assert [n for n in names if "_Thr_" in n] == ["i1_Thr_1", "i1_Thr_2", "i2_Thr_1"]
    th = m.thresholds["i1"]["values"]
    assert th[0] == pytest.approx(-th[1]) and th[1] == pytest.approx(0.6091404, abs=1e-6)


def test_endogenous_covariates(pheno):
    cont = pheno.drop(columns=["i1", "i2"])
    m = buildOneFac(cont, [f"i{k}" for k in range(3, 8)], covariates=["covar1"], exogenous=False)
    assert "covar1" in m.manifestVars and "covar1_to_F" in m.parameters()[0]
    # all continuous + no exogenous covariates -> cumulants: the mean structure is dropped (R/model.R:497-503)
    assert "covar1_var" in m.parameters()[0] and "covar1_mean" not in m.parameters()[0]
    # a factor column anywhere in the observed data has typeof() == 'integer' in R: the reference warns and stays
    # on marginals (R/model.R:486-496), even when the model does not use that column
    with pytest.warns(UserWarning, match="Cumulants method prevented by integer observed columns"):
        m = buildOneFac(pheno, [f"i{k}" for k in range(3, 8)], covariates=["covar1"], exogenous=False)
    assert m.continuousType == "marginals" and "covar1_mean" in m.parameters()[0]
    m = buildOneFac(pheno, [f"i{k}" for k in range(1, 8)], covariates=["covar1"], exogenous=False)
    assert m.continuousType == "marginals" and "covar1_mean" in m.parameters()[0]


def test_two_fac_and_gxe(pheno):
    m = buildTwoFac(pheno, ["i3", "i4", "i5"], ["i5", "i6", "i7"])
    names, start, _, _ = m.parameters()
    assert "facCov" in names and start[names.index("facCov")] == .3
    assert {"snp_to_F1", "snp_to_F2", "F1_lambda_i5", "F2_lambda_i5"} <= set(names)
    g = buildItem(pheno, "i3", gxe=["covar1", "i4"])
    assert g.manifestVars == ["i3", "snp", "snp_covar1", "snp_i4"]        # test-gxe.R:35-38
    assert {"snp_covar1_to_i3", "snp_i4_to_i3", "snp_to_i3"} <= set(g.parameters()[0])


def test_flatten_roundtrip(pheno):
    m = buildOneFacRes(pheno[[f"i{k}" for k in range(3, 8)]], [f"i{k}" for k in range(3, 8)])
    f = flatten(m)
    assert f["nm"] == 6 and f["nv"] == 7 and f["continuous_type"] == "cumulants"
    assert len(f["par_names"]) == 16 and f["cells"].shape[1] == 5
    free = f["cells"][f["cells"][:, 3] >= 0]
    assert set(free[:, 3].astype(int)) == set(range(16))


def test_prepare_compute_plan_and_namespace(pheno):
    """Every export of the reference NAMESPACE exists; prepareComputePlan attaches the per-SNP loop (R/model.R:122-195)."""
    import gwsem_b200 as g
    for name in ["GWAS", "buildAnalysesPlan", "buildItem", "buildOneFac", "buildOneFacRes", "buildOneItem", "buildTwoFac",
                 "isSuspicious", "loadResults", "loadSuspicious", "numAvailableRecords", "prepareComputePlan",
                 "setupExogenousCovariates", "setupThresholds", "signif", "signifGxE"]:
        assert callable(getattr(g, name)), name
    m = g.buildItem(pheno, "i1")
    with pytest.raises(ValueError, match="Rejected"):
        g.prepareComputePlan(m, "x.pgen", "out.log", "stray")
    m = g.prepareComputePlan(m, "x.pgen", "o.log", SNP=[1, 2], startFrom=1)
    text = repr(m.compute)
    assert "mxComputeLoadData('x.pgen'" in text and "mxComputeCheckpoint('o.log'" in text and "StandardError" in text
    assert "NumericDeriv" not in text
    ml = g.prepareComputePlan(g.buildItem(pheno, "i1", fitfun="ML"), "x.bgen")
    assert "NumericDeriv" in repr(ml.compute) and "HessianQuality" in repr(ml.compute)
    with pytest.raises(RuntimeError, match="defunct"):
        g.loadSuspicious("o.log", "snp_to_i1")
