"""The reference tests' simulated phenotypes, replayed with the bit-exact R RNG of oracle/r_rng.py.

Each function repeats the preamble of one reference test file (file:line in its docstring) on
tests/golden/extdata/example.psam and returns the same data frame the R code ends up with, so
that the numeric expectations written in those tests become known-answer tests here.
"""
import os

import numpy as np
import pandas as pd

from r_rng import RRng, cut_codes, quantile7

EXTDATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extdata")


def read_psam():
    ph = pd.read_csv(os.path.join(EXTDATA, "example.psam"), sep="\t")
    ph = ph.rename(columns={ph.columns[0]: "FID"})
    return ph


def _indicators(rng, pheno, num):
    """loadings <- rbeta(num, 4, 3); resid <- rbeta(num, 4, 3)^2;
    indicators <- pheno$phenotype %*% t(loadings) + mvrnorm(nrow(pheno), rep(0, num), diag(num))"""
    loadings = rng.rbeta(num, 4, 3)
    rng.rbeta(num, 4, 3)   # `resid`: drawn, never used
    ind = np.outer(pheno["phenotype"].to_numpy(float), loadings) + rng.mvrnorm_identity(len(pheno), num)
    return loadings, ind


def test_model_pheno():
    """tests/testthat/test-model.R:10-30: seven indicators, i1 cut at its tertiles, i2 at its median.
    Returns (pheno, origPheno, loadings)."""
    rng = RRng(1)
    pheno = read_psam()
    loadings, ind = _indicators(rng, pheno, 7)
    for j in range(7):
        pheno[f"i{j + 1}"] = ind[:, j]
    orig = pheno.copy()
    b1 = np.concatenate([[-np.inf], quantile7(ind[:, 0], [1 / 3, 2 / 3]), [np.inf]])
    b2 = np.concatenate([[-np.inf], quantile7(ind[:, 1], [0.5]), [np.inf]])
    pheno["i1"] = pd.Categorical(cut_codes(ind[:, 0], b1), categories=[0, 1, 2], ordered=True)
    pheno["i2"] = pd.Categorical(cut_codes(ind[:, 1], b2), categories=[0, 1], ordered=True)
    return pheno, orig, loadings


def test_formats_pheno():
    """tests/testthat/test-formats.R:7-22: five continuous indicators."""
    rng = RRng(1)
    pheno = read_psam()
    loadings, ind = _indicators(rng, pheno, 5)
    for j in range(5):
        pheno[f"i{j + 1}"] = ind[:, j]
    return pheno, loadings


def test_gxe_pheno():
    """tests/testthat/test-gxe.R:7-26: seven continuous indicators, column 3 renamed 'sex'."""
    rng = RRng(1)
    pheno = read_psam().rename(columns={"SEX": "sex"})
    loadings, ind = _indicators(rng, pheno, 7)
    for j in range(7):
        pheno[f"i{j + 1}"] = ind[:, j]
    return pheno, loadings


def signif(est, var):
    """R/report.R:183-185: Z = est / sqrt(V), P = 2 pnorm(-|Z|)."""
    from scipy.special import ndtr
    z = np.asarray(est) / np.sqrt(np.asarray(var))
    return z, 2 * ndtr(-np.abs(z))
