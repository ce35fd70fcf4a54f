"""Writes src/check/example_call.txt: the .Call("gwsem_run", ...) arguments of the reference's C1 case
(tests/testthat/test-covariate.R:34-39: buildOneFac, 2 ordinal + 5 continuous items, 2 exogenous covariates on
example.pgen) plus a rowFilter, as R/flatten.R assembles them.  Run from the repository root."""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np

import ref_pins as rp
from gwsem_b200.gwas import call_arguments, dump_call
from gwsem_b200.model import buildOneFac

warnings.simplefilter("ignore")
pheno, _, _ = rp.test_model_pheno()
rng = np.random.default_rng(5)
pheno["covar1"], pheno["covar2"] = rng.normal(size=500), rng.normal(size=500)
skip = np.zeros(500, bool)
skip[::9] = True
model = buildOneFac(pheno[~skip].reset_index(drop=True), [f"i{k}" for k in range(1, 8)], covariates=["covar1", "covar2"])
args = call_arguments(model, "../tests/golden/extdata/example.pgen", "/tmp/gwsem_check_out.log", SNP=[3, 4, 5, 120],
                      rowFilter=skip)
dump_call(args, os.path.join(ROOT, "src", "check", "example_call.txt"))
