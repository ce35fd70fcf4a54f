"""The .Call layer (src/gwsem_call.c) without R: the arguments R/flatten.R assembles (here produced by the Python
mirror, gwsem_b200.gwas.call_arguments) go through the real shim compiled against stand-in R headers (src/stub/).
CPU: against a mock of the C ABI that checks every field (make -C src check).  GPU: against libgwsem_b200.so,
and the log must equal the one GWAS() writes for the same call."""
import os
import subprocess
import warnings

import numpy as np
import pandas as pd
import pytest

import ref_pins as rp
from conftest import EXTDATA, ROOT
from gwsem_b200.gwas import call_arguments, dump_call
from gwsem_b200.model import buildItem, buildOneFac, buildTwoFac

SRC = os.path.join(ROOT, "src")


def models():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pheno, _, _ = rp.test_model_pheno()
        rng = np.random.default_rng(5)
        pheno["covar1"], pheno["covar2"] = rng.normal(size=500), rng.normal(size=500)
        skip = np.zeros(500, bool)
        skip[::7] = True
        kept = pheno[~skip].reset_index(drop=True)
        return {
            "item_gxe_cumulants": (buildItem(pheno[["i3", "i4"]], "i3", gxe=["i4"]), dict(SNP=[1, 2, 3, 28])),
            "onefac_marginals_rowfilter": (buildOneFac(kept, [f"i{k}" for k in range(1, 8)], covariates=["covar1", "covar2"]),
                                           dict(SNP=[3, 4, 5], rowFilter=skip)),
            "item_two_depvars_exogenous_snp": (buildItem(pheno, ["i2", "i3"]), dict(startFrom=195)),
            "twofac_ml": (buildTwoFac(pheno[[f"i{k}" for k in range(3, 8)]], ["i3", "i4", "i5"], ["i5", "i6", "i7"], fitfun="ML"),
                          dict(SNP=[7, 8])),
        }


def build(target):
    subprocess.run(["make", "-C", SRC, target], check=True, capture_output=True)


def test_make_check_with_the_committed_example():
    out = subprocess.run(["make", "-C", SRC, "check"], capture_output=True, text=True)
    assert out.returncode == 0 and "check_shim: ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("name", ["item_gxe_cumulants", "onefac_marginals_rowfilter", "item_two_depvars_exogenous_snp", "twofac_ml"])
def test_marshalling_against_the_mock(tmp_path, name):
    build("check/check_shim_mock")
    model, kw = models()[name]
    dump = str(tmp_path / "call.txt")
    dump_call(call_arguments(model, os.path.join(EXTDATA, "example.pgen"), str(tmp_path / "out.log"), **kw), dump)
    out = subprocess.run([os.path.join(SRC, "check", "check_shim_mock"), dump, "mock"], capture_output=True, text=True)
    assert out.returncode == 0 and "check_shim: ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["item_gxe_cumulants", "onefac_marginals_rowfilter", "item_two_depvars_exogenous_snp"])
def test_dot_call_layer_runs_the_same_gwas(engine, tmp_path, name):
    from gwsem_b200 import GWAS
    build("check/check_shim_real")
    model, kw = models()[name]
    snp = os.path.join(EXTDATA, "example.pgen")
    out_c, out_py = str(tmp_path / "c.log"), str(tmp_path / "py.log")
    dump = str(tmp_path / "call.txt")
    dump_call(call_arguments(model, snp, out_c, **kw), dump)
    run = subprocess.run([os.path.join(SRC, "check", "check_shim_real"), dump, "real"], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    res = GWAS(model, snp, out_py, **kw)
    assert f"rows written: {res.rows_written}" in run.stdout
    a = pd.read_csv(out_c, sep="\t", quoting=3).drop(columns=["timestamp"])
    b = pd.read_csv(out_py, sep="\t", quoting=3).drop(columns=["timestamp"])
    pd.testing.assert_frame_equal(a, b, check_exact=True)
