"""GPU parity of the ML / FIML path (the C5 shape: buildTwoFac with a
gxe term, fitfun='ML') against oracle/ml_oracle.py, through the C ABI: hardcalls (with missing
calls: those people keep contributing through their items), PLINK 1 coding, dosages, and the
closed form for buildItem (ML regression of the item on the genotype)."""
import numpy as np
import pandas as pd
import pytest

from gwsem_b200 import synth
from gwsem_b200.gwas import model_columns
from gwsem_b200.model import buildItem, buildOneFac, buildOneFacRes, buildTwoFac, flatten

pytestmark = pytest.mark.gpu
EST_RTOL, SE_RTOL = 1e-4, 1e-3


def cohort(n, seed, n_cov=1):
    rng = np.random.default_rng(seed)
    F1 = rng.normal(size=n)
    F2 = 0.3 * F1 + rng.normal(size=n)
    ph = pd.DataFrame({f"pc{k + 1}": rng.normal(size=n) for k in range(n_cov)})
    ph["E"] = rng.integers(0, 2, n).astype(float)
    for k, (f, lam) in enumerate([(F1, .8), (F1, .7), (F1, .6), (F2, .6), (F2, .9), (F2, .5)], 1):
        ph[f"i{k}"] = lam * f + sum(0.2 * ph[f"pc{c + 1}"] for c in range(n_cov)) + rng.normal(size=n) + 0.5 * k
    return ph


def compare(res, want, names):
    from gwsem_b200.engine import STATUS_TEXT
    for i, w in enumerate(want):
        if w["catch1"]:
            assert res.catch_code[i] != 0, (i, w["catch1"])
            continue
        assert res.catch_code[i] == 0, (i, res.catch_code[i])
        assert STATUS_TEXT[int(res.status[i])] == w["status"] == "OK", (i, res.status[i], w["status"])
        assert res.n_used[i] == w["n"]
        se_w, se_g = np.sqrt(np.diag(w["vcov"])), np.sqrt(np.diag(res.vcov[i]))
        tol = EST_RTOL * np.maximum(np.abs(w["est"]), se_w)
        bad = np.abs(res.est[i] - w["est"]) > tol
        assert not bad.any(), (i, [names[k] for k in np.nonzero(bad)[0]], res.est[i][bad], w["est"][bad])
        assert np.allclose(se_g, se_w, rtol=SE_RTOL, atol=0), (i, np.max(np.abs(se_g / se_w - 1)))
        assert res.fit[i] == pytest.approx(w["fit"], rel=1e-9)


def build(case, ph):
    if case == "twofac_gxe_cov":
        return buildTwoFac(ph, ["i1", "i2", "i3"], ["i4", "i5", "i6"], gxe="E", covariates=["pc1"], fitfun="ML")
    if case == "onefac":
        return buildOneFac(ph, ["i1", "i2", "i3", "i4"], fitfun="ML")
    if case == "onefacres_cov":
        return buildOneFacRes(ph, ["i1", "i2", "i3"], covariates=["pc1", "pc2"], fitfun="ML")
    if case == "item_cov":
        return buildItem(ph, "i1", covariates=["pc1"], fitfun="ML")
    raise KeyError(case)


@pytest.mark.parametrize("case", ["twofac_gxe_cov", "onefac", "onefacres_cov", "item_cov"])
def test_ml_matches_oracle(engine, case):
    import ml_oracle as mo
    from gwsem_b200.engine import Model
    n = 1200
    ph = cohort(n, 17, n_cov=2)
    model = build(case, ph)
    flat = flatten(model)
    assert flat["fitfun"] == "ML"
    data, ncat = model_columns(model)
    m_snps = 5
    codes = synth.simulate_hardcalls(n, m_snps, seed=3, maf_range=(0.15, 0.5))
    codes[1, :50] = 3            # missing calls
    codes[2] = 0                 # monomorphic -> minVariance catch
    codes[4, ::9] = 3
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    want = [mo.fit_snp_ml(flat, data, g) for g in geno]
    assert [bool(w["catch1"]) for w in want] == [False, False, True, False, False]
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    compare(res, want, flat["par_names"])
    bed = np.array([3, 2, 0, 1], np.uint8)[codes]
    res2 = gm.fit_2bit(synth.pack_2bit(bed), plink1=True)
    ok = res.catch_code == 0
    assert np.array_equal(res2.est[ok].view(np.uint64), res.est[ok].view(np.uint64))
    gm.close()


def test_ml_item_is_ml_regression(engine):
    """buildItem under ML regresses the item on the genotype (a definition variable): slope and
    intercept are OLS, the residual variance is the ML one (/ n), and 2 H^-1 gives the ML standard
    error sqrt(s2 / Sxx) of the slope."""
    from gwsem_b200.engine import Model
    n = 3000
    rng = np.random.default_rng(2)
    codes = synth.simulate_hardcalls(n, 4, seed=8, maf_range=(0.2, 0.5))
    codes[3, :100] = 3
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    ph = pd.DataFrame({"y": 0.15 * np.nan_to_num(geno[0]) + rng.normal(size=n) + 3.0})
    model = buildItem(ph, "y", fitfun="ML")
    flat = flatten(model)
    assert flat["manifest"] == ["y"] and flat["exo_pred"] == ["snp"]
    data, ncat = model_columns(model)
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    names = flat["par_names"]
    for i, g in enumerate(geno):
        ok = np.isfinite(g)
        X = np.column_stack([np.ones(ok.sum()), g[ok]])
        beta, *_ = np.linalg.lstsq(X, ph["y"][ok], rcond=None)
        r = ph["y"][ok] - X @ beta
        s2 = r @ r / ok.sum()
        assert res.status[i] == 0 and res.n_used[i] == ok.sum()
        assert res.est[i][names.index("snp_to_y")] == pytest.approx(beta[1], rel=1e-7, abs=1e-9)
        assert res.est[i][names.index("yMean")] == pytest.approx(beta[0], rel=1e-7)
        assert res.est[i][names.index("y_res")] == pytest.approx(s2, rel=1e-7)
        k = names.index("snp_to_y")
        assert res.vcov[i][k, k] == pytest.approx(s2 * np.linalg.inv(X.T @ X)[1, 1], rel=1e-5)
    gm.close()


def test_ml_dosage_rows_and_gwas_log(engine, tmp_path):
    """Dosages (.pgen with dosage tracks, the generic double-column path) through GWAS(): the log
    carries fitUnits -2lnL and the estimates match the oracle."""
    import os
    import ml_oracle as mo
    from conftest import EXTDATA, GOLDEN
    from gwsem_b200 import GWAS
    ps = pd.read_csv(os.path.join(EXTDATA, "example.psam"), sep="\t")
    rng = np.random.default_rng(5)
    ph = pd.DataFrame({"pc1": rng.normal(size=500), "E": rng.integers(0, 2, 500).astype(float)})
    for k in range(1, 5):
        ph[f"i{k}"] = ps["phenotype"] * (0.4 + 0.1 * k) + 0.1 * ph["pc1"] + rng.normal(size=500)
    model = buildOneFac(ph, ["i1", "i2", "i3", "i4"], gxe="E", covariates=["pc1"], fitfun="ML")
    flat = flatten(model)
    data, ncat = model_columns(model)
    pick = [1, 2, 77, 199]
    for fname, gold_name in (("example.pgen", "example_pgen.npz"), ("example.bed", "example_bed.npz")):
        out = str(tmp_path / "out.log")
        GWAS(model, os.path.join(EXTDATA, fname), out, SNP=pick)
        log = pd.read_csv(out, sep="\t", quoting=3, dtype={"catch1": str, "statusCode": str})
        gold = np.load(os.path.join(GOLDEN, gold_name))
        assert list(log["fitUnits"]) == ["-2lnL"] * len(pick)
        for row, snp in zip(log.to_dict("records"), [s - 1 for s in pick]):
            w = mo.fit_snp_ml(flat, data, gold["geno"][snp])
            if w["catch1"]:
                assert isinstance(row["catch1"], str) and row["catch1"]
                continue
            assert row["statusCode"] == "OK"
            se = np.sqrt(np.diag(w["vcov"]))
            for k, nme in enumerate(flat["par_names"]):
                assert abs(row[nme] - w["est"][k]) <= EST_RTOL * max(abs(w["est"][k]), se[k]), (fname, snp, nme)
                assert row[f"V{nme}:{nme}"] == pytest.approx(w["vcov"][k, k], rel=2 * SE_RTOL), (fname, snp, nme)
