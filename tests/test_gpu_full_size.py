"""Size-independent properties at the cohort sizes BASELINE.json quotes (C2: 100k people, C3: 50k people), where
the numpy oracles only afford a couple of SNPs:
  * C2 (buildItem, one continuous phenotype, WLS cumulants) is a saturated model with a closed form
    (SURVEY.md section 8a row O3): snp_to_y = S_ys / S_ss, snp_res = S_ss, y_res = S_yy - S_ys^2 / S_ss;
  * a row never depends on its neighbours (the reference restarts every SNP from the original starts,
    R/model.R:188): any sub-batch, in any order, gives the same bits;
  * PLINK 1 coding of the same genotypes gives the same bits;
  * C3 (buildOneFac, 3 ordinal items + 5 PCs, WLS marginals): 32 SNPs -- null, causal (|rho| > 0.04: the exact Newton
    passes) and with 0.5 % missing calls (the exact re-estimation kernel) -- against oracle/marginals_oracle.py within
    the north-star tolerances (the oracle runs in a process pool: ~5 s per SNP at 50k people), everything else through
    the two invariances;
  * C2 estimates *and* their ADF covariance at 100k people against oracle/fit_oracle.py;
  * C4 (buildOneFacRes, 5 continuous items, .bed coding, 50k people) against oracle/fit_oracle.py;
  * C5 (buildTwoFac ML, 8 items + gxe, bgen dosages, 200k people) through GWAS() against oracle/ml_oracle.py."""
import os

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


def _oracle_marginals_job(args):
    import warnings
    warnings.simplefilter("ignore")
    import marginals_oracle as mo
    flat, data, ncat, g = args
    w = mo.fit_snp_marginals(flat, data, g, ncat)
    return {k: w[k] for k in ("est", "vcov", "status", "catch1", "n")}


def compare_rows(res, want, names, idx=None):
    from gwsem_b200.engine import STATUS_TEXT
    for k, w in enumerate(want):
        i = k if idx is None else idx[k]
        assert not w["catch1"] and res.catch_code[i] == 0, (i, w["catch1"])
        assert STATUS_TEXT[int(res.status[i])] == w["status"] and res.n_used[i] == w["n"], i
        se_w, se_g = np.sqrt(np.diag(w["vcov"])), np.sqrt(np.diag(res.vcov[i]))
        bad = np.abs(res.est[i] - w["est"]) > EST_RTOL * np.maximum(np.abs(w["est"]), se_w)
        assert not bad.any(), (i, [names[j] for j in np.nonzero(bad)[0]], res.est[i][bad], w["est"][bad])
        assert np.allclose(se_g, se_w, rtol=SE_RTOL, atol=0), (i, np.max(np.abs(se_g / se_w - 1)))


def test_c3_32_snps_null_causal_and_missing_against_the_oracle(engine):
    """The headline configuration at full size, the branches the bench times: null SNPs (series root accepted),
    causal SNPs (exact Newton passes), SNPs with 0.5 % missing calls (marg_exact_kernel)."""
    import multiprocessing as mp
    from gwsem_b200.engine import Model
    n, m = 50_000, 32
    rng = np.random.default_rng(77)
    codes = synth.simulate_hardcalls(n, m, seed=1003, maf_range=(0.05, 0.5))
    causal = [2, 9, 17, 30]
    ph = c3_cohort(n)
    # the items are re-cut from a factor that carries the causal SNPs, so that |rho(snp, item)| > 0.04 for them
    K = 5
    X = np.column_stack([ph[f"pc{k + 1}"] for k in range(K)])
    g = codes[causal].astype(float)
    gz = (g - g.mean(1, keepdims=True)) / g.std(1, keepdims=True)
    F = rng.normal(size=n) + X @ np.full(K, 0.1) + 0.13 * gz.sum(0)
    cuts = [[-.43, .43], [-.67, 0, .67], [-.84, -.25, .25, .84]]
    for j, lam in enumerate([.8, .7, .6]):
        ph[f"i{j + 1}"] = pd.Categorical(np.searchsorted(cuts[j], lam * F + rng.normal(size=n)), ordered=True)
    missing = [4, 9, 12, 21, 28]                    # one of them causal
    for r in missing:
        codes[r][rng.random(n) < 0.005] = 3
    model = buildOneFac(ph, ["i1", "i2", "i3"], covariates=[f"pc{k + 1}" for k in range(K)])
    flat = flatten(model)
    data, ncat = model_columns(model)
    gm = Model(engine, flat, data, ncat)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    assert (res.status == 0).all() and (res.catch_code == 0).all()
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    with mp.get_context("spawn").Pool(min(os.cpu_count() or 1, 16)) as pool:
        want = pool.map(_oracle_marginals_job, [(flat, data, ncat, geno[i]) for i in range(m)])
    names = flat["par_names"]
    k = names.index("snp_to_F")
    z = np.array([w["est"][k] / np.sqrt(w["vcov"][k, k]) for w in want])
    assert (np.abs(z[causal]) > 8).all() and np.median(np.abs(np.delete(z, causal))) < 1.5     # the causal ones are far out
    assert all(want[r]["n"] < n for r in missing)
    compare_rows(res, want, names)
    gm.close()


def test_c2_covariance_at_100k_people(engine):
    """BASELINE.json configs[1]: the ADF covariance users read (Z = est / sqrt(V)) at full size."""
    import fit_oracle as fo
    from gwsem_b200.engine import Model
    n, m = 100_000, 12
    rng = np.random.default_rng(2001)
    codes = synth.simulate_hardcalls(n, m, seed=1001)
    codes[5][rng.random(n) < 0.005] = 3
    y = 0.05 * np.where(codes[3] == 3, 0, codes[3]) + rng.standard_t(6, size=n)     # heavy tails: ADF differs from OLS
    model = buildItem(pd.DataFrame({"y": y}), "y")
    flat = flatten(model)
    gm = Model(engine, flat, {"y": y})
    res = gm.fit_2bit(synth.pack_2bit(codes))
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    want = [fo.fit_snp_cumulants(flat, {"y": y}, g) for g in geno]
    compare_rows(res, want, flat["par_names"])
    for i, w in enumerate(want):          # every covariance entry, not only the diagonal
        assert np.allclose(res.vcov[i], w["vcov"], rtol=2 * SE_RTOL, atol=1e-12 * np.abs(w["vcov"]).max()), i
    gm.close()


def test_c4_residuals_model_on_bed_coding_at_50k_people(engine):
    """BASELINE.json configs[3]: buildOneFacRes, 5 continuous items, PLINK 1 coding."""
    import fit_oracle as fo
    from gwsem_b200.engine import Model
    from gwsem_b200.model import buildOneFacRes
    n, m = 50_000, 8
    rng = np.random.default_rng(2003)
    codes = synth.simulate_hardcalls(n, m, seed=1003)
    codes[2][rng.random(n) < 0.01] = 3
    F = rng.normal(size=n)
    ph = pd.DataFrame({f"i{k + 1}": lam * F + rng.normal(size=n) + 0.03 * (codes[1] % 3) for k, lam in enumerate([.8, .7, .6, .5, .4])})
    model = buildOneFacRes(ph, list(ph.columns))
    flat = flatten(model)
    assert flat["continuous_type"] == "cumulants"
    data, ncat = model_columns(model)
    gm = Model(engine, flat, data, ncat)
    bed = np.array([3, 2, 0, 1], np.uint8)[codes]
    res = gm.fit_2bit(synth.pack_2bit(bed), plink1=True)
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    want = [fo.fit_snp_cumulants(flat, data, g) for g in geno]
    compare_rows(res, want, flat["par_names"])
    gm.close()


def test_c5_two_factor_ml_on_bgen_dosages_at_200k_people(engine, tmp_path):
    """BASELINE.json configs[4]: buildTwoFac ML, 8 items + a gxe term, bgen v1.2 layout 2 with 8-bit probabilities,
    200k people, through GWAS() (file -> host inflate -> device bit-field decode -> dosage moments -> ML fit -> log)."""
    import ml_oracle
    from gwsem_b200 import GWAS
    from gwsem_b200.model import buildTwoFac
    n, m = 200_000, 6
    rng = np.random.default_rng(2004)
    codes = synth.simulate_hardcalls(n, m, seed=1004, maf_range=(0.1, 0.5))
    # probabilities around the hardcall: 8-bit triples summing to 255
    p = np.zeros((m, n, 3), np.int64)
    noise = rng.integers(1, 40, size=(m, n))
    other = (codes + 1 + rng.integers(0, 2, size=(m, n))) % 3
    np.put_along_axis(p, (2 - codes)[..., None].astype(np.int64), (255 - noise)[..., None], axis=2)     # P(AA) first: A1 dosage
    np.put_along_axis(p, (2 - other)[..., None].astype(np.int64), noise[..., None], axis=2)
    path = str(tmp_path / "c5.bgen")
    synth.write_bgen(path, p[..., :2], layout=2, bits=8, compression=1)
    F1 = rng.normal(size=n)
    F2 = 0.3 * F1 + rng.normal(size=n)
    ph = pd.DataFrame({"E": rng.integers(0, 2, n).astype(float)})
    for k in range(8):
        ph[f"i{k + 1}"] = (0.5 + 0.05 * k) * (F1 if k < 4 else F2) + rng.normal(size=n)
    model = buildTwoFac(ph, [f"i{k}" for k in range(1, 5)], [f"i{k}" for k in range(5, 9)], gxe="E", fitfun="ML")
    flat = flatten(model)
    data, ncat = model_columns(model)
    out = str(tmp_path / "c5.log")
    GWAS(model, path, out)
    log = pd.read_csv(out, sep="\t", quoting=3, dtype={"catch1": str, "statusCode": str})
    assert len(log) == m
    from gwsem_b200.engine import Source
    with Source(path, "bgen", n) as src:
        geno, _ = src.load_rows(engine, list(range(m)))
    for row, gi in zip(log.to_dict("records"), geno):
        w = ml_oracle.fit_snp_ml(flat, data, gi)
        assert row["statusCode"] == w["status"] == "OK"
        se = np.sqrt(np.diag(w["vcov"]))
        for k, nme in enumerate(flat["par_names"]):
            assert abs(row[nme] - w["est"][k]) <= EST_RTOL * max(abs(w["est"][k]), se[k]), (row["MxComputeLoop1"], nme)
            assert np.sqrt(row[f"V{nme}:{nme}"]) == pytest.approx(se[k], rel=SE_RTOL), (row["MxComputeLoop1"], nme)
