"""GPU parity of the WLS marginals path (ordinal items and / or exogenous covariates, the C3
shape: buildOneFac with ordinal items + PC covariates) against oracle/marginals_oracle.py,
through the C ABI, on hardcalls, on PLINK 1 coding and on dosages."""
import numpy as np
import pandas as pd
import pytest

from gwsem_b200 import synth
from gwsem_b200.gwas import model_columns
from gwsem_b200.model import buildOneFac, buildOneFacRes, buildTwoFac, flatten

pytestmark = pytest.mark.gpu
EST_RTOL, SE_RTOL = 1e-4, 1e-3


def cohort(n, seed, n_items, n_cov, ordinal):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, n_cov))
    f = rng.normal(size=n)
    ph = pd.DataFrame({f"pc{k + 1}": X[:, k] for k in range(n_cov)})
    cuts = [[-.5, .5], [-.8, 0, .8], [-1, -.3, .3, 1.], [0.], [-.4, .6]]
    for j in range(n_items):
        ystar = (0.8 - 0.1 * j) * f + X @ (rng.normal(size=n_cov) * 0.2) + rng.normal(size=n)
        if ordinal[j]:
            ph[f"i{j + 1}"] = pd.Categorical(np.searchsorted(cuts[j % len(cuts)], ystar), ordered=True)
        else:
            ph[f"i{j + 1}"] = ystar + 0.3
    return ph


def compare(res, want, names):
    from gwsem_b200.engine import STATUS_TEXT
    for i, w in enumerate(want):
        if w["catch1"]:
            assert res.catch_code[i] != 0, (i, w["catch1"])
            continue
        assert res.catch_code[i] == 0, (i, res.catch_code[i])
        assert STATUS_TEXT[int(res.status[i])] == w["status"], (i, res.status[i], w["status"])
        assert res.n_used[i] == w["n"]
        se_w, se_g = np.sqrt(np.diag(w["vcov"])), np.sqrt(np.diag(res.vcov[i]))
        tol = EST_RTOL * np.maximum(np.abs(w["est"]), se_w)
        bad = np.abs(res.est[i] - w["est"]) > tol
        assert not bad.any(), (i, [names[k] for k in np.nonzero(bad)[0]], res.est[i][bad], w["est"][bad])
        assert np.allclose(se_g, se_w, rtol=SE_RTOL, atol=0), (i, np.max(np.abs(se_g / se_w - 1)))


CASES = {
    "c3_ordinal_pcs": dict(n_items=3, n_cov=3, ordinal=[True, True, True], build="onefac"),
    "mixed_items": dict(n_items=4, n_cov=2, ordinal=[True, False, True, False], build="onefac"),
    "continuous_with_covariates": dict(n_items=4, n_cov=2, ordinal=[False] * 4, build="onefacres"),
    "ordinal_no_covariates": dict(n_items=3, n_cov=0, ordinal=[True, True, True], build="onefac"),
    "twofac_ordinal": dict(n_items=6, n_cov=1, ordinal=[True, False, True, True, False, True], build="twofac"),
    # reference tests/testthat/test-covariate.R:68-75: exogenous = FALSE turns the covariates into manifest
    # variables, so 2 ordinal + 5 continuous items + 2 covariates are nine "items" besides snp
    "endogenous_covariates": dict(n_items=7, n_cov=2, ordinal=[True, True] + [False] * 5, build="onefac_endo"),
}


@pytest.mark.parametrize("case", list(CASES))
def test_marginals_matches_oracle(engine, case):
    import marginals_oracle as mo
    from gwsem_b200.engine import Model
    cfg = CASES[case]
    n = 1500
    ph = cohort(n, 11, cfg["n_items"], cfg["n_cov"], cfg["ordinal"])
    items = [f"i{j + 1}" for j in range(cfg["n_items"])]
    covs = [f"pc{k + 1}" for k in range(cfg["n_cov"])] or None
    if cfg["build"] == "onefac":
        model = buildOneFac(ph, items, covariates=covs)
    elif cfg["build"] == "onefac_endo":
        model = buildOneFac(ph, items, covariates=covs, exogenous=False)
    elif cfg["build"] == "onefacres":
        model = buildOneFacRes(ph, items, covariates=covs)
    else:
        model = buildTwoFac(ph, items[:3], items[3:], covariates=covs)
    flat = flatten(model)
    assert flat["continuous_type"] == "marginals"
    data, ncat = model_columns(model)
    m_snps = 6
    codes = synth.simulate_hardcalls(n, m_snps, seed=5, maf_range=(0.15, 0.5))
    codes[2] = 0                                   # monomorphic -> minVariance catch
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    want = [mo.fit_snp_marginals(flat, data, g, ncat) for g in geno]
    assert sum(w["status"] == "OK" for w in want) >= m_snps - 1
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    compare(res, want, flat["par_names"])
    bed = np.array([3, 2, 0, 1], np.uint8)[codes]
    res2 = gm.fit_2bit(synth.pack_2bit(bed), plink1=True)
    assert np.array_equal(res2.est.view(np.uint64), res.est.view(np.uint64))
    gm.close()


@pytest.mark.parametrize("rate,n", [(0.01, 4000), (0.04, 600), (0.3, 900)])
def test_marginals_with_missing_calls(engine, rate, n):
    """Missing calls drop people from every statistic of that SNP (naAction='omit', R/model.R:435-436):
    marg_exact_kernel re-estimates every stage-1 / stage-2 quantity and every score product over the
    retained people, so the north-star tolerances hold at any missing rate (1 %, the 4 % of the
    reference's example.bed, 30 %)."""
    import marginals_oracle as mo
    from gwsem_b200.engine import Model
    ph = cohort(n, 21, 3, 3, [True, False, True])
    items = ["i1", "i2", "i3"]
    model = buildOneFac(ph, items, covariates=["pc1", "pc2", "pc3"])
    flat = flatten(model)
    data, ncat = model_columns(model)
    codes = synth.simulate_hardcalls(n, 5, seed=9, maf_range=(0.2, 0.5), missing_rate=rate)
    codes[3] = np.where(codes[3] == 3, 0, codes[3])          # one call-complete SNP in the same batch
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    want = [mo.fit_snp_marginals(flat, data, g, ncat) for g in geno]
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    compare(res, want, flat["par_names"])
    bed = np.array([3, 2, 0, 1], np.uint8)[codes]
    res2 = gm.fit_2bit(synth.pack_2bit(bed), plink1=True)
    assert np.array_equal(res2.est.view(np.uint64), res.est.view(np.uint64))
    # a row does not depend on its neighbours: any sub-batch, any order, the same bits
    res3 = gm.fit_2bit(synth.pack_2bit(codes[[4, 0]]))
    assert np.array_equal(res3.est.view(np.uint64), res.est[[4, 0]].view(np.uint64))
    gm.close()


def test_large_batches_are_repeatable_bit_for_bit(engine):
    """compute-sanitizer is closed on this pool, so data races have to show up as run-to-run differences: a batch
    large enough to keep every SM full (the multi-warp fit kernel once raced in its pivot search and only showed
    it beyond ~8,000 SNPs per launch), three times, must give identical bits -- and the same bits as small batches."""
    from gwsem_b200.engine import Model
    n, m = 1500, 12_000
    ph = cohort(n, 17, 3, 2, [True, False, True])
    model = buildOneFac(ph, ["i1", "i2", "i3"], covariates=["pc1", "pc2"])
    flat = flatten(model)
    data, ncat = model_columns(model)
    codes = synth.simulate_hardcalls(n, m, seed=23, maf_range=(0.05, 0.5))
    codes[::97, ::11] = 3                      # some SNPs through the exact re-estimation kernel
    packed = synth.pack_2bit(codes)
    gm = Model(engine, flat, data, ncat)
    runs = [gm.fit_2bit(packed) for _ in range(3)]
    for r in runs[1:]:
        for f in ("est", "vcov", "fit"):
            assert np.array_equal(getattr(r, f).view(np.uint64), getattr(runs[0], f).view(np.uint64)), f
        assert np.array_equal(r.status, runs[0].status) and np.array_equal(r.catch_code, runs[0].catch_code)
    small = gm.fit_2bit(packed[5000:5300])
    assert np.array_equal(small.est.view(np.uint64), runs[0].est[5000:5300].view(np.uint64))
    assert (runs[0].status[runs[0].catch_code == 0] == 0).mean() > 0.99
    gm.close()


def test_items_with_exogenous_snp_match_oracle(engine):
    """buildItem with several depVars, one of them ordinal (tests/testthat/test-model.R:86): exogenous defaults
    to TRUE, snp is a definition variable of every item and of their residual correlation."""
    import marginals_oracle as mo
    from gwsem_b200.engine import Model
    from gwsem_b200.model import buildItem
    n = 1200
    ph = cohort(n, 31, 3, 2, [True, False, True])
    model = buildItem(ph, ["i1", "i2", "i3"], covariates=["pc1", "pc2"])
    flat = flatten(model)
    assert flat["exo_pred"][0] == "snp" and flat["manifest"] == ["i1", "i2", "i3"]
    data, ncat = model_columns(model)
    codes = synth.simulate_hardcalls(n, 5, seed=3, maf_range=(0.2, 0.5))
    codes[1, ::9] = 3
    codes[2] = 0
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    want = [mo.fit_snp_marginals(flat, data, g, ncat) for g in geno]
    assert [bool(w["catch1"]) for w in want] == [False, False, True, False, False]
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    compare(res, want, flat["par_names"])
    gm.close()


def test_gxe_product_as_manifest_variable_matches_oracle(engine):
    """A factor model with ordinal items and a gxe term: `snp_<moderator>` is a manifest variable filled per SNP
    with snp * moderator (R/model.R:427-434, tests/testthat/test-gxe.R:67-77)."""
    import marginals_oracle as mo
    from gwsem_b200.engine import Model
    n = 1500
    ph = cohort(n, 41, 3, 1, [True, False, True])
    ph["mod"] = np.random.default_rng(2).normal(size=n) + 1.0
    model = buildOneFac(ph, ["i1", "i2", "i3"], covariates=["pc1"], gxe="mod")
    flat = flatten(model)
    assert flat["manifest"] == ["snp", "i1", "i2", "i3", "snp_mod"]
    data, ncat = model_columns(model)
    codes = synth.simulate_hardcalls(n, 4, seed=6, maf_range=(0.25, 0.5))
    codes[1, ::13] = 3
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    want = [mo.fit_snp_marginals(flat, data, g, ncat) for g in geno]
    assert all(w["status"] == "OK" for w in want)
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    compare(res, want, flat["par_names"])
    gm.close()


def test_gwas_ordinal_one_factor_on_fixtures(engine, tmp_path):
    """C1 shape: buildOneFac with ordinal items and covariates on the bundled example files, through
    GWAS(): hardcalls (.bed) and dosages (.pgen, the generic double-column path)."""
    import os
    import marginals_oracle as mo
    from conftest import EXTDATA, GOLDEN
    from gwsem_b200 import GWAS
    ps = pd.read_csv(os.path.join(EXTDATA, "example.psam"), sep="\t")
    rng = np.random.default_rng(1)
    ph = pd.DataFrame({"covar1": rng.normal(size=500), "covar2": rng.normal(size=500)})
    for k in range(1, 5):
        ph[f"i{k}"] = ps["phenotype"] * (0.4 + 0.1 * k) + rng.normal(size=500)
    ph["i1"] = pd.Categorical(pd.qcut(ph["i1"], 3, labels=False), ordered=True)
    ph["i2"] = pd.Categorical(pd.qcut(ph["i2"], 2, labels=False), ordered=True)
    items = ["i1", "i2", "i3", "i4"]
    model = buildOneFac(ph, items, covariates=["covar1", "covar2"])
    flat = flatten(model)
    data, ncat = model_columns(model)
    for fname, gold_name in (("example.bed", "example_bed.npz"), ("example.pgen", "example_pgen.npz")):
        out = str(tmp_path / "out.log")
        GWAS(model, os.path.join(EXTDATA, fname), out, SNP=[1, 2, 3, 50, 120])
        log = pd.read_csv(out, sep="\t", quoting=3, dtype={"catch1": str, "statusCode": str})
        gold = np.load(os.path.join(GOLDEN, gold_name))
        for row, snp in zip(log.to_dict("records"), [0, 1, 2, 49, 119]):
            w = mo.fit_snp_marginals(flat, data, gold["geno"][snp], ncat)
            if w["catch1"]:
                assert isinstance(row["catch1"], str) and row["catch1"]
                continue
            assert row["statusCode"] == w["status"]
            se = np.sqrt(np.diag(w["vcov"]))
            n_miss = int(np.isnan(gold["geno"][snp]).sum())
            est_tol, var_tol = 1e-4, 2e-3     # with or without missing calls (example.bed: ~4 % of 500 people)
            for k, nme in enumerate(flat["par_names"]):
                assert abs(row[nme] - w["est"][k]) <= est_tol * max(abs(w["est"][k]), se[k]), (fname, snp, nme, n_miss)
                assert row[f"V{nme}:{nme}"] == pytest.approx(w["vcov"][k, k], rel=var_tol), (fname, snp, nme, n_miss)


def item_cohort(n, seed, cuts, n_cov):
    rng = np.random.default_rng(seed)
    ph = pd.DataFrame({"sex": rng.integers(0, 2, n).astype(float)})
    for k in range(1, n_cov):
        ph[f"pc{k}"] = rng.normal(size=n)
    ystar = rng.normal(size=n) + 0.3 * ph["sex"].to_numpy() + sum(0.2 * ph[c].to_numpy() for c in ph.columns[1:])
    ph["anx"] = pd.Categorical(np.searchsorted(cuts, ystar), ordered=True)
    return ph


@pytest.mark.parametrize("cuts,n_cov", [([-0.5, 0.6], 2), ([0.2], 0), ([-1.0, -0.2, 0.4, 1.1], 3)])
def test_build_item_ordinal_matches_oracle(engine, cuts, n_cov):
    """buildItem with an ordered phenotype (the README example, case/control when there are two
    levels): snp is an exogenous predictor, so every SNP re-fits the ordered probit of the item on
    snp + covariates.  People with a missing call are dropped exactly (no down-date here)."""
    import marginals_oracle as mo
    from gwsem_b200.engine import Model
    from gwsem_b200.model import buildItem
    n = 1500
    ph = item_cohort(n, 3, cuts, n_cov)
    covs = [c for c in ph.columns if c != "anx"][:n_cov] or None
    model = buildItem(ph[["anx"] + (covs or [])], "anx", covariates=covs)
    flat = flatten(model)
    assert flat["continuous_type"] == "marginals" and flat["manifest"] == ["anx"] and flat["exo_pred"][0] == "snp"
    data, ncat = model_columns(model)
    codes = synth.simulate_hardcalls(n, 6, seed=5, maf_range=(0.15, 0.5))
    codes[1, :60] = 3                              # missing calls
    codes[2] = 0                                   # monomorphic: the probit on a constant snp is not identified
    codes[4, ::7] = 3
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    want = [mo.fit_snp_marginals(flat, data, g, ncat) for g in geno]
    assert [bool(w["catch1"]) for w in want] == [False, False, True, False, False, False]
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    compare(res, want, flat["par_names"])
    assert np.allclose(res.maf[[0, 3, 5]], [min(g.mean() / 2, 1 - g.mean() / 2) for g in geno[[0, 3, 5]]])
    bed = np.array([3, 2, 0, 1], np.uint8)[codes]
    res2 = gm.fit_2bit(synth.pack_2bit(bed), plink1=True)
    ok = res.catch_code == 0
    assert np.array_equal(res2.est[ok].view(np.uint64), res.est[ok].view(np.uint64))
    gm.close()


def test_build_item_ordinal_gwas_on_fixtures(engine, tmp_path):
    """The README's first example through GWAS(): hardcalls (.bed) and dosages (.pgen)."""
    import os
    import marginals_oracle as mo
    from conftest import EXTDATA, GOLDEN
    from gwsem_b200 import GWAS, loadResults, signif
    from gwsem_b200.model import buildItem
    ps = pd.read_csv(os.path.join(EXTDATA, "example.psam"), sep="\t")
    rng = np.random.default_rng(4)
    ph = pd.DataFrame({"sex": rng.integers(0, 2, 500).astype(float)})
    ph["anxiety"] = pd.Categorical(pd.qcut(ps["phenotype"] + rng.normal(size=500), 3, labels=False), ordered=True)
    model = buildItem(ph, "anxiety", covariates=["sex"])
    flat = flatten(model)
    data, ncat = model_columns(model)
    pick = [1, 2, 3, 50, 120, 199]
    for fname, gold_name in (("example.bed", "example_bed.npz"), ("example.pgen", "example_pgen.npz")):
        out = str(tmp_path / "out.log")
        GWAS(model, os.path.join(EXTDATA, fname), out, SNP=pick)
        log = pd.read_csv(out, sep="\t", quoting=3, dtype={"catch1": str, "statusCode": str})
        gold = np.load(os.path.join(GOLDEN, gold_name))
        assert len(log) == len(pick)
        for row, snp in zip(log.to_dict("records"), [s - 1 for s in pick]):
            w = mo.fit_snp_marginals(flat, data, gold["geno"][snp], ncat)
            if w["catch1"]:
                assert isinstance(row["catch1"], str) and row["catch1"]
                continue
            assert row["statusCode"] == w["status"]
            se = np.sqrt(np.diag(w["vcov"]))
            for k, nme in enumerate(flat["par_names"]):
                assert abs(row[nme] - w["est"][k]) <= EST_RTOL * max(abs(w["est"][k]), se[k]), (fname, snp, nme)
                assert row[f"V{nme}:{nme}"] == pytest.approx(w["vcov"][k, k], rel=2 * SE_RTOL), (fname, snp, nme)
        lr = signif(loadResults(out, "snp_to_anxiety"), "snp_to_anxiety")
        assert np.isfinite(lr["P"]).sum() >= len(pick) - 1


@pytest.mark.parametrize("gxe,n_cov", [("mod", 2), (None, 1), ("mod", 0)])
def test_build_item_continuous_exogenous_matches_oracle(engine, gxe, n_cov):
    """buildItem(..., fitfun='WLS', exogenous=TRUE) with a continuous depVar: the gene-environment
    vignette's model (tools/vignettes/GeneEnvironmentInteraction.Rmd:47-49).  snp and snp_<moderator>
    are definition variables; per SNP this is the regression of the item on them and the covariates."""
    import marginals_oracle as mo
    from gwsem_b200.engine import Model
    from gwsem_b200.model import buildItem
    n = 1500
    rng = np.random.default_rng(8)
    codes = synth.simulate_hardcalls(n, 6, seed=12, maf_range=(0.15, 0.5))
    codes[1, :70] = 3
    codes[2] = 0                                   # monomorphic
    codes[5, ::11] = 3
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    ph = pd.DataFrame({"mod": rng.normal(size=n)})
    for k in range(n_cov):
        ph[f"pc{k + 1}"] = rng.normal(size=n)
    ph["phe"] = (0.2 * np.nan_to_num(geno[0]) + 0.3 * ph["mod"] + 0.15 * np.nan_to_num(geno[0]) * ph["mod"]
                 + sum(0.1 * ph[f"pc{k + 1}"] for k in range(n_cov)) + rng.normal(size=n) + 2.0)
    covs = (["mod"] if gxe else []) + [f"pc{k + 1}" for k in range(n_cov)]
    model = buildItem(ph, "phe", covariates=covs or None, fitfun="WLS", exogenous=True, gxe=gxe)
    flat = flatten(model)
    assert flat["manifest"] == ["phe"] and flat["exo_pred"][0] == "snp" and flat["continuous_type"] == "marginals"
    data, ncat = model_columns(model)
    want = [mo.fit_snp_marginals(flat, data, g, ncat) for g in geno]
    assert [bool(w["catch1"]) for w in want] == [False, False, True, False, False, False]
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    compare(res, want, flat["par_names"])
    assert res.catch_code[2] == 1                  # observed variance of snp below minVariance
    # closed form: OLS coefficients
    names = flat["par_names"]
    ok = np.isfinite(geno[0])
    cols = [np.ones(ok.sum()), geno[0][ok]] + ([geno[0][ok] * ph["mod"][ok]] if gxe else []) + [ph[c][ok] for c in covs]
    beta = np.linalg.lstsq(np.column_stack(cols), ph["phe"][ok], rcond=None)[0]
    assert res.est[0][names.index("snp_to_phe")] == pytest.approx(beta[1], rel=1e-9)
    if gxe:
        assert res.est[0][names.index("snp_mod_to_phe")] == pytest.approx(beta[2], rel=1e-9)
    bed = np.array([3, 2, 0, 1], np.uint8)[codes]
    res2 = gm.fit_2bit(synth.pack_2bit(bed), plink1=True)
    good = res.catch_code == 0
    assert np.array_equal(res2.est[good].view(np.uint64), res.est[good].view(np.uint64))
    gm.close()


def test_gxe_vignette_model_through_gwas(engine, tmp_path):
    """The vignette's GxE call through GWAS() on the bundled files (hardcalls and dosages), with the
    follow-up the vignette does: loadResults + signifGxE on the interaction term."""
    import os
    import marginals_oracle as mo
    from conftest import EXTDATA, GOLDEN
    from gwsem_b200 import GWAS, loadResults, signifGxE
    from gwsem_b200.model import buildItem
    ps = pd.read_csv(os.path.join(EXTDATA, "example.psam"), sep="\t")
    rng = np.random.default_rng(6)
    ph = pd.DataFrame({"mod": rng.normal(size=500), "pc1": rng.normal(size=500)})
    ph["phe"] = ps["phenotype"] + 0.3 * ph["mod"] + rng.normal(size=500)
    model = buildItem(ph, "phe", covariates=["mod", "pc1"], fitfun="WLS", exogenous=True, gxe="mod")
    flat = flatten(model)
    data, ncat = model_columns(model)
    pick = [1, 2, 90, 199]
    for fname, gold_name in (("example.bed", "example_bed.npz"), ("example.pgen", "example_pgen.npz")):
        out = str(tmp_path / "out.log")
        GWAS(model, os.path.join(EXTDATA, fname), out, SNP=pick)
        log = pd.read_csv(out, sep="\t", quoting=3, dtype={"catch1": str, "statusCode": str})
        gold = np.load(os.path.join(GOLDEN, gold_name))
        for row, snp in zip(log.to_dict("records"), [s - 1 for s in pick]):
            w = mo.fit_snp_marginals(flat, data, gold["geno"][snp], ncat)
            if w["catch1"]:
                assert isinstance(row["catch1"], str) and row["catch1"]
                continue
            assert row["statusCode"] == "OK"
            se = np.sqrt(np.diag(w["vcov"]))
            for k, nme in enumerate(flat["par_names"]):
                assert abs(row[nme] - w["est"][k]) <= EST_RTOL * max(abs(w["est"][k]), se[k]), (fname, snp, nme)
                assert row[f"V{nme}:{nme}"] == pytest.approx(w["vcov"][k, k], rel=2 * SE_RTOL), (fname, snp, nme)
        lr = signifGxE(loadResults(out, "snp_mod_to_phe"), "snp_mod_to_phe", level=0.5)
        assert np.isfinite(lr["P"]).sum() >= len(pick) - 1


def test_build_item_ordinal_with_gxe_matches_oracle(engine):
    """An ordered depVar with a gxe term (the vignette: "It is possible to conduct GxE analyses on binary or
    ordinal variables"): snp and snp_<moderator> are definition variables of the ordered probit."""
    import marginals_oracle as mo
    from gwsem_b200.engine import Model
    from gwsem_b200.model import buildItem
    n = 2000
    rng = np.random.default_rng(9)
    codes = synth.simulate_hardcalls(n, 4, seed=14, maf_range=(0.2, 0.5))
    codes[1, :50] = 3
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    ph = pd.DataFrame({"mod": rng.integers(0, 2, n).astype(float), "pc1": rng.normal(size=n)})
    ystar = 0.2 * np.nan_to_num(geno[0]) + 0.3 * ph["mod"] + 0.2 * np.nan_to_num(geno[0]) * ph["mod"] + 0.1 * ph["pc1"] + rng.normal(size=n)
    ph["phe"] = pd.Categorical(np.searchsorted([0.2, 1.1], ystar), ordered=True)
    model = buildItem(ph, "phe", covariates=["mod", "pc1"], gxe="mod")
    flat = flatten(model)
    assert flat["exo_pred"][:2] == ["snp", "snp_mod"] and flat["continuous_type"] == "marginals"
    data, ncat = model_columns(model)
    want = [mo.fit_snp_marginals(flat, data, g, ncat) for g in geno]
    assert all(w["status"] == "OK" for w in want)
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    compare(res, want, flat["par_names"])
    gm.close()


def _c3_like(n, seed):
    rng = np.random.default_rng(seed)
    K = 5
    X = rng.normal(size=(n, K))
    causal = synth.simulate_hardcalls(n, 2, seed=seed + 1, maf_range=(0.2, 0.4))
    gz = (causal - causal.mean(1, keepdims=True)) / causal.std(1, keepdims=True)
    F = rng.normal(size=n) + X @ np.full(K, 0.1) + 0.15 * gz.sum(0)
    ph = pd.DataFrame({f"pc{k + 1}": X[:, k] for k in range(K)})
    cuts = [[-.43, .43], [-.67, 0, .67], [-.84, -.25, .25, .84]]
    for j, lam in enumerate([.8, .7, .6]):
        ystar = lam * F + X @ (rng.normal(size=K) * 0.1) + rng.normal(size=n)
        ph[f"i{j + 1}"] = pd.Categorical(np.searchsorted(cuts[j], ystar), ordered=True)
    return ph, causal


def test_series_path_and_structured_fit_match_the_per_person_kernels(engine, monkeypatch):
    """The class-sum (series) statistics + the structured one-warp fit against the per-person statistics kernel + the
    generic fit on the same batch: null SNPs (series path), causal SNPs (per-person kernel started from the series root),
    rare alleles, a monomorphic row.  Tolerances far inside the oracle tolerances; any sub-batch gives the same bits."""
    from gwsem_b200.engine import Model
    n, m = 20000, 300
    ph, causal = _c3_like(n, 77)
    model = buildOneFac(ph, ["i1", "i2", "i3"], covariates=[f"pc{k + 1}" for k in range(5)])
    flat = flatten(model)
    data, ncat = model_columns(model)
    codes = synth.simulate_hardcalls(n, m, seed=78, maf_range=(0.05, 0.5))
    codes[10:40] = synth.simulate_hardcalls(n, 30, seed=79, maf_range=(0.008, 0.04))   # rare alleles: large |u| in the hom class
    codes[50] = causal[0]
    codes[51] = causal[1]
    codes[60] = 0
    packed = synth.pack_2bit(codes)
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(packed)
    sub = gm.fit_2bit(synth.pack_2bit(codes[[51, 7, 299, 20]]))
    assert np.array_equal(sub.est.view(np.uint64), res.est[[51, 7, 299, 20]].view(np.uint64))
    assert np.array_equal(sub.vcov.view(np.uint64), res.vcov[[51, 7, 299, 20]].view(np.uint64))
    gm.close()
    monkeypatch.setenv("GWSEM_NO_SERIES", "1")
    monkeypatch.setenv("GWSEM_FIT_GENERIC", "1")
    gm2 = Model(engine, flat, data, ncat)
    ref = gm2.fit_2bit(packed)
    gm2.close()
    assert np.array_equal(res.status, ref.status) and np.array_equal(res.catch_code, ref.catch_code)
    assert res.catch_code[60] != 0 and (res.catch_code[40:60] == 0).all() and (res.catch_code[61:] == 0).all()
    assert (res.catch_code[10:40] == 0).sum() >= 20   # (the rarest fall under minVariance, in both paths)
    ok = res.catch_code == 0
    se = np.sqrt(np.diagonal(ref.vcov[ok], axis1=1, axis2=2))
    se2 = np.sqrt(np.diagonal(res.vcov[ok], axis1=1, axis2=2))
    d = np.abs(res.est[ok] - ref.est[ok]) / np.maximum(np.abs(ref.est[ok]), se)
    assert d.max() < 2e-5, d.max()
    assert np.max(np.abs(se2 / se - 1)) < 1e-5
    k = flat["par_names"].index("snp_to_F")
    z = res.est[:, k] / np.sqrt(res.vcov[:, k, k])
    assert abs(z[50]) > 8 and abs(z[51]) > 8                      # the causal SNPs are found


def test_series_path_with_row_filter_and_missing_phenotypes(engine, monkeypatch):
    """rowFilter (people skipped in the genotype file) and phenotype NAs (naAction = 'omit') under the class-sum
    statistics path: excluded people carry zero features, the class counts come from the masked popcounts.  Against
    the oracle on the retained people, and against the per-person kernels."""
    import marginals_oracle as mo
    from gwsem_b200.engine import Model
    n_src = 9000
    rng = np.random.default_rng(5)
    rf = rng.random(n_src) < 0.2
    n = int((~rf).sum())
    ph = cohort(n, 41, 3, 2, [True, True, True])
    ph.loc[ph.index[::29], "pc1"] = np.nan          # naAction = 'omit' on a covariate
    ph["i2"] = ph["i2"].astype(object)
    ph.loc[ph.index[5::31], "i2"] = np.nan          # ... and on an ordinal item
    ph["i2"] = pd.Categorical(ph["i2"], ordered=True)
    model = buildOneFac(ph, ["i1", "i2", "i3"], covariates=["pc1", "pc2"])
    flat = flatten(model)
    data, ncat = model_columns(model)
    codes = synth.simulate_hardcalls(n_src, 24, seed=6, maf_range=(0.1, 0.5))
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes[:, ~rf]]
    want = [mo.fit_snp_marginals(flat, data, g, ncat) for g in geno[:6]]
    gm = Model(engine, flat, data, ncat)
    gm.set_row_filter(rf)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    gm.close()
    compare_first = type("R", (), {})()
    for f in ("est", "vcov", "status", "catch_code", "n_used"):
        setattr(compare_first, f, getattr(res, f)[:6])
    compare(compare_first, want, flat["par_names"])
    monkeypatch.setenv("GWSEM_NO_SERIES", "1")
    monkeypatch.setenv("GWSEM_FIT_GENERIC", "1")
    gm2 = Model(engine, flat, data, ncat)
    gm2.set_row_filter(rf)
    ref = gm2.fit_2bit(synth.pack_2bit(codes))
    gm2.close()
    assert np.array_equal(res.status, ref.status) and np.array_equal(res.n_used, ref.n_used)
    se = np.sqrt(np.diagonal(ref.vcov, axis1=1, axis2=2))
    assert np.max(np.abs(res.est - ref.est) / np.maximum(np.abs(ref.est), se)) < 2e-5
    assert np.max(np.abs(np.sqrt(np.diagonal(res.vcov, axis1=1, axis2=2)) / se - 1)) < 1e-5


def test_series_path_binary_and_seven_level_items(engine, monkeypatch):
    """Items with two, four and seven categories, no covariates: the class-sum path against the per-person kernels."""
    from gwsem_b200.engine import Model
    n = 15000
    rng = np.random.default_rng(91)
    F = rng.normal(size=n)
    ph = pd.DataFrame()
    for j, (lam, cuts) in enumerate([(.8, [0.2]), (.7, [-.67, 0, .67]), (.6, [-1.2, -.6, -.2, .2, .6, 1.2])]):
        ph[f"i{j + 1}"] = pd.Categorical(np.searchsorted(cuts, lam * F + rng.normal(size=n)), ordered=True)
    model = buildOneFac(ph, ["i1", "i2", "i3"])
    flat = flatten(model)
    data, ncat = model_columns(model)
    assert ncat == {"i1": 2, "i2": 4, "i3": 7}
    packed = synth.pack_2bit(synth.simulate_hardcalls(n, 64, seed=92, maf_range=(0.03, 0.5)))
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(packed)
    gm.close()
    monkeypatch.setenv("GWSEM_NO_SERIES", "1")
    monkeypatch.setenv("GWSEM_FIT_GENERIC", "1")
    gm2 = Model(engine, flat, data, ncat)
    ref = gm2.fit_2bit(packed)
    gm2.close()
    assert (ref.status == 0).all() and np.array_equal(res.status, ref.status)
    se = np.sqrt(np.diagonal(ref.vcov, axis1=1, axis2=2))
    assert np.max(np.abs(res.est - ref.est) / np.maximum(np.abs(ref.est), se)) < 2e-5
    assert np.max(np.abs(np.sqrt(np.diagonal(res.vcov, axis1=1, axis2=2)) / se - 1)) < 1e-5
