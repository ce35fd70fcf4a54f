"""The CUDA path against the reference's OWN numeric expectations (known-answer tests).

The same assertions as tests/test_reference_pins.py (which pins the CPU oracles), made on what
GWAS() -> gwsem_gwas_run (C ABI) -> checkpoint log -> loadResults / signif produces on the bundled
fixture files, with the phenotypes the reference's tests simulate (tests/ref_pins.py replays R's RNG).
Every test also checks the log against the oracle at the north-star tolerances (estimates 1e-4, SEs 1e-3).
"""
import os
import warnings

import numpy as np
import pandas as pd
import pytest

import ref_pins as rp
from conftest import EXTDATA, GOLDEN
from test_reference_pins import ITEMS, all_equal, column, fivenum, run

pytestmark = pytest.mark.gpu
EST_RTOL, SE_RTOL = 1e-4, 1e-3


def gwas(model, fname, tmp_path, **kw):
    from gwsem_b200 import GWAS
    out = str(tmp_path / "out.log")
    res = GWAS(model, os.path.join(EXTDATA, fname), out, **kw)
    log = pd.read_csv(out, sep="\t", quoting=3, dtype={"catch1": str, "statusCode": str})
    log["catch1"] = log["catch1"].fillna("")
    return out, log, res


def check_against_oracle(log, flat, rows):
    assert len(log) == len(rows)
    for row, w in zip(log.to_dict("records"), rows):
        if w["catch1"]:
            assert row["catch1"], (row["MxComputeLoop1"], w["catch1"])
            continue
        assert row["catch1"] == "" and row["statusCode"] == w["status"], (row["MxComputeLoop1"], row["catch1"], row["statusCode"])
        se = np.sqrt(np.diag(w["vcov"]))
        for k, nme in enumerate(flat["par_names"]):
            assert abs(row[nme] - w["est"][k]) <= EST_RTOL * max(abs(w["est"][k]), se[k]), (row["MxComputeLoop1"], nme)
            assert np.sqrt(row[f"V{nme}:{nme}"]) == pytest.approx(se[k], rel=SE_RTOL), (row["MxComputeLoop1"], nme)


def geno(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))["geno"]


@pytest.fixture(scope="module")
def model_pheno():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return rp.test_model_pheno()


def test_build_item_wls_and_ml_pins(engine, tmp_path, model_pheno):
    """tests/testthat/test-model.R:49-79"""
    from gwsem_b200 import loadResults, signif
    from gwsem_b200.model import buildItem
    pheno = model_pheno[0]
    m = buildItem(pheno, "i3")
    out, log, _ = gwas(m, "example.pgen", tmp_path, SNP=[3, 4, 5])
    pgen = signif(loadResults(out, "snp_to_i3"), "snp_to_i3")
    assert list(pgen["SNP"]) == ["RSID_4", "RSID_5", "RSID_6"]
    assert int(np.argmin(pgen["P"])) + 1 == 3
    assert all_equal(pgen["P"][2], 0.155, 1e-2)
    check_against_oracle(log, *run(m, geno("example_pgen"), [3, 4, 5], "cumulants"))
    with pytest.raises(Exception, match="out of data"):          # test-model.R:60-66
        gwas(m, "example.pgen", tmp_path, SNP=[250])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        oi = buildItem(pheno, "i3", fitfun="ML", minMAF=0.1)
    out, log, _ = gwas(oi, "example.pgen", tmp_path, SNP=[3, 4, 5])
    ml = signif(loadResults(out, "snp_to_i3"), "snp_to_i3")
    assert abs(np.corrcoef(ml["P"], pgen["P"])[0, 1] - 1) <= 1e-3
    check_against_oracle(log, *run(oi, geno("example_pgen"), [3, 4, 5], "ml"))


def test_build_item_two_dep_vars_pins(engine, tmp_path, model_pheno):
    """tests/testthat/test-model.R:86-103: an ordinal and a continuous depVar, snp exogenous."""
    from gwsem_b200 import loadResults, signif
    from gwsem_b200.model import buildItem
    m = buildItem(model_pheno[0], ["i2", "i3"])
    out, log, _ = gwas(m, "example.pgen", tmp_path, SNP=list(range(11, 21)))
    for focus, rx, p in (("snp_to_i2", 10, 0.0251), ("snp_to_i3", 4, 0.086)):
        r = signif(loadResults(out, focus), focus)
        assert int(np.argmin(r["P"])) + 1 == rx and all_equal(r["P"].min(), p, 1e-2)
    check_against_oracle(log, *run(m, geno("example_pgen"), range(11, 21), "marginals"))


def test_one_factor_pins(engine, tmp_path, model_pheno):
    """tests/testthat/test-model.R:111-128"""
    from gwsem_b200 import loadResults, signif
    from gwsem_b200.model import buildOneFac
    pheno, _, loadings = model_pheno
    m = buildOneFac(pheno, ITEMS)
    out, log, _ = gwas(m, "example.pgen", tmp_path, SNP=[3, 4, 5])
    pgen = signif(loadResults(out, ["snp_to_F", "lambda_i1"]), "snp_to_F", signAdj="lambda_i1")
    pgen2 = signif(loadResults([out, out], ["snp_to_F", "lambda_i1"]), "snp_to_F", signAdj="lambda_i1")
    assert list(pgen["SNP"]) == [f"RSID_{k}" for k in (4, 5, 6)] and list(pgen2["SNP"]) == 2 * list(pgen["SNP"])
    assert np.array_equal(np.round(pgen["snp_to_F"], 3), [0.242, 0.072, -0.087])   # see test_reference_pins.py
    assert all_equal(pgen["snp_to_F"], [0.242, 0.072, -0.087], 2.5e-3)
    assert all_equal(log[["lambda_" + i for i in ITEMS]].mean(axis=0) / loadings, np.ones(7), 0.2)
    check_against_oracle(log, *run(m, geno("example_pgen"), [3, 4, 5], "marginals"))


def test_one_factor_residuals_pins(engine, tmp_path, model_pheno):
    """tests/testthat/test-model.R:137-157"""
    from gwsem_b200 import loadResults, signif
    from gwsem_b200.model import buildOneFacRes
    pheno, orig, loadings = model_pheno
    m2 = buildOneFacRes(pheno, ITEMS)
    iF = m2.vars.index("F")
    for i in ITEMS:
        m2.A["lbound"][m2.vars.index(i), iF] = 0
    out, log, _ = gwas(m2, "example.pgen", tmp_path, SNP=[3, 4, 5])
    q13 = np.repeat(rp.quantile7(orig["i1"], [1 / 3]), 3)
    assert np.mean(np.abs(log["i1_Thr_1"] - q13)) / np.mean(np.abs(log["i1_Thr_1"])) <= 0.21   # see test_reference_pins.py
    assert np.mean(np.abs(log["i2_Thr_1"])) <= 0.15
    assert all_equal(log[["lambda_" + i for i in ITEMS]].mean(axis=0) / loadings, np.ones(7), 0.2)
    r = signif(loadResults(out, "snp_to_i2"), "snp_to_i2")
    assert all_equal(r["P"], [0.455, 0.885, 0.395], 1e-2)
    check_against_oracle(log, *run(m2, geno("example_pgen"), [3, 4, 5], "marginals"))


def test_two_factor_pins(engine, tmp_path, model_pheno):
    """tests/testthat/test-model.R:161-171"""
    from gwsem_b200.model import buildTwoFac
    pheno, _, loadings = model_pheno
    m = buildTwoFac(pheno, ITEMS[:4], ITEMS[3:])
    out, log, _ = gwas(m, "example.pgen", tmp_path, SNP=[3, 4, 5])
    names = [f"F1_lambda_i{k}" for k in (1, 2, 3)] + [f"F2_lambda_i{k}" for k in (5, 6, 7)]
    assert all_equal(log[names].mean(axis=0) / np.delete(loadings, 3), np.ones(6), 0.15)
    assert all_equal(log["facCov"], np.repeat(1.17, 3), 0.02)
    check_against_oracle(log, *run(m, geno("example_pgen"), [3, 4, 5], "marginals"))


def test_gxe_moderator_level_pins(engine, tmp_path):
    """tests/testthat/test-gxe.R:51-65"""
    from gwsem_b200 import isSuspicious, loadResults, signifGxE
    from gwsem_b200.model import buildItem
    pheno, _ = rp.test_gxe_pheno()
    m = buildItem(pheno, "i3", gxe=["i4"])
    out, log, _ = gwas(m, "example.pgen", tmp_path)
    m1 = loadResults(out, "snp_i4_to_i3")
    d = (signifGxE(m1, "snp_i4_to_i3", level=0.5)["Z"] - signifGxE(m1, "snp_i4_to_i3", level=0.7)["Z"]).to_numpy()
    assert all_equal(fivenum(d[~np.isnan(d)]), [-0.55, -0.34, -0.27, -0.19, 0.04], 1e-2)
    bad = isSuspicious(m1, ["snp_to_i3", "snp_i4_to_i3"])
    assert [int((~bad).sum()), int(bad.sum())] == [197, 2]
    check_against_oracle(log, *run(m, geno("example_pgen"), range(1, 200), "cumulants"))


@pytest.fixture(scope="module")
def formats_model():
    from gwsem_b200.model import buildOneFac
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pheno, _ = rp.test_formats_pheno()
        return buildOneFac(pheno, [f"i{k}" for k in range(1, 6)])


def test_formats_bed_pins(engine, tmp_path, formats_model):
    """tests/testthat/test-formats.R:53-70: hardcalls with 4 % missing calls (naAction = 'omit')."""
    from gwsem_b200 import isSuspicious, loadResults, signif
    out, log, res = gwas(formats_model, "example.bed", tmp_path)
    assert len(log) == 199 and res.loadCounter == 1
    lr = signif(loadResults(out, ["snp_to_F", "lambda_i1"]), "snp_to_F", signAdj="lambda_i1")
    bad = isSuspicious(lr)
    assert np.max(np.abs(np.array([(~bad).sum(), bad.sum()]) - [173, 26])) <= 6
    assert all_equal(fivenum(lr[~bad]["Z"]), [-2.61, -0.6, -0.03, 0.6, 2.61], 1e-2)
    check_against_oracle(log, *run(formats_model, geno("example_bed"), range(1, 200), "marginals"))


def test_formats_pgen_and_bgen_pins(engine, tmp_path, formats_model):
    """tests/testthat/test-formats.R:35-51, 72-127"""
    from gwsem_b200 import loadResults
    out, pgen, res = gwas(formats_model, "example.pgen", tmp_path)
    assert len(pgen) == 199 and res.loadCounter == 1
    rng = (np.nanmin(res.last_snp), np.nanmax(res.last_snp))
    assert abs(rng[0]) <= 0.02 and abs(rng[1] - 2) <= 0.02
    check_against_oracle(pgen, *run(formats_model, geno("example_pgen"), range(1, 200), "marginals"))
    out, bgen, res = gwas(formats_model, "example.bgen", tmp_path)
    assert len(bgen) == 199 and res.loadCounter == 1
    assert list(loadResults(out, "snp_to_F").columns) == ["MxComputeLoop1", "CHR", "BP", "SNP", "A1", "A2", "statusCode",
                                                          "catch1", "snp_to_F", "Vsnp_to_F:snp_to_F"]
    out, last, _ = gwas(formats_model, "example.bgen", tmp_path, startFrom=190)
    assert len(last) == 10 and list(last["SNP"]) == list(bgen["SNP"][189:199])
    assert list(pgen["A1"]) == list(pgen["A1"]) and set(bgen["A1"]) == set(pgen["A1"])
    bgen = bgen.set_index(bgen["SNP"].str.replace("^SNP", "RS", regex=True)).loc[pgen["SNP"]].reset_index(drop=True)
    assert all_equal(bgen["MAF"], pgen["MAF"], 1e-2)
    mask = (bgen["catch1"] == "") & (pgen["catch1"] == "") & (bgen["statusCode"] == "OK") & (pgen["statusCode"] == "OK")
    assert abs(int(mask.sum()) - 197) <= 2
    assert np.sqrt(np.mean((bgen["snp_to_F"][mask] - pgen["snp_to_F"][mask]) ** 2)) <= 1.8
