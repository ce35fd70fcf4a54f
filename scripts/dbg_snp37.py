"""Development aid: one SNP of the test-formats.R bed run, fast path vs exact path vs oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import warnings
import numpy as np
warnings.simplefilter("ignore")
import ref_pins as rp
import marginals_oracle as mo
from gwsem_b200 import synth
from gwsem_b200.engine import Engine, Model
from gwsem_b200.gwas import model_columns
from gwsem_b200.model import buildOneFac, flatten
pheno, _ = rp.test_formats_pheno()
m = buildOneFac(pheno, [f"i{k}" for k in range(1, 6)])
flat = flatten(m); data, ncat = model_columns(m)
G = np.load(os.path.join(ROOT, "tests/golden/example_bed.npz"))["geno"]
pick = [36, 0, 5]
codes = np.where(np.isnan(G[pick]), 3, G[pick]).astype(np.uint8)
eng = Engine(0)
gm = Model(eng, flat, data, ncat)
res = gm.fit_2bit(synth.pack_2bit(codes))
for i, s in enumerate(pick):
    w = mo.fit_snp_marginals(flat, data, G[s], ncat)
    se = np.sqrt(np.diag(w["vcov"]))
    d = (res.est[i] - w["est"]) / np.maximum(np.abs(w["est"]), se)
    print(s, "n", res.n_used[i], w["n"], "max rel diff", np.abs(d).max(), flat["par_names"][int(np.abs(d).argmax())],
          "se rel", np.abs(np.sqrt(np.diag(res.vcov[i])) / se - 1).max(), "fit", res.fit[i], w["fit"], "it", res.iterations[i], w["iterations"])
