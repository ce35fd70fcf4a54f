"""GPU parity of the hardcall hot path, through the C ABI:
packed 2-bit records -> (a) decoded double columns, bit-exact against the reference
(golden fixtures from oracle/_ref) and the C oracle; (b) per-SNP WLS fits against
oracle/fit_oracle.py within the tolerances BASELINE.json states (estimates 1e-4
relative, standard errors 1e-3 relative, identical status)."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import EXTDATA, GOLDEN
from gwsem_b200 import synth
from gwsem_b200.model import buildItem, buildOneFac, buildOneFacRes, buildTwoFac, flatten

pytestmark = pytest.mark.gpu

EST_RTOL, SE_RTOL = 1e-4, 1e-3   # BASELINE.json north_star tolerances


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def lut_decode(codes):
    lut = np.array([0.0, 1.0, 2.0, np.nan])
    out = lut[codes]
    out.view(np.uint64)[np.isnan(out)] = 0x7FF00000000007A2
    return out


def bed_rows(name, n):
    raw = np.fromfile(os.path.join(EXTDATA, name), np.uint8)[3:]
    return raw.reshape(-1, (n + 3) // 4)


@pytest.mark.parametrize("name,n", [("example.bed", 500), ("group1.bed", 100), ("group2.bed", 100)])
def test_decode_bed_fixture_matches_reference(engine, name, n):
    gold = np.load(os.path.join(GOLDEN, name.replace(".", "_") + ".npz"))
    rows = bed_rows(name, n)
    geno, maf = engine.decode_2bit(rows, n, plink1=True)
    assert bits_equal(geno, gold["geno"])
    assert bits_equal(maf, gold["maf"])
    geno, maf = engine.decode_2bit(rows, n, plink1=True, row_filter=gold["row_filter"])
    assert bits_equal(geno, gold["geno_filtered"])
    assert bits_equal(maf, gold["maf_filtered"])


def test_decode_test2_pgen_fixture_matches_reference(engine):
    gold = np.load(os.path.join(GOLDEN, "test2_pgen.npz"))
    raw = np.fromfile(os.path.join(EXTDATA, "test2.pgen"), np.uint8)
    rows = raw[-50:].reshape(10, 5)  # ten vrtype-0 records of ceil(20/4) bytes close the file
    geno, maf = engine.decode_2bit(rows, 20)
    assert bits_equal(geno, gold["geno"])
    assert bits_equal(maf, gold["maf"])
    geno, maf = engine.decode_2bit(rows, 20, row_filter=gold["row_filter"])
    assert bits_equal(geno, gold["geno_filtered"])


@pytest.mark.parametrize("n,m", [(1, 3), (31, 5), (4097, 64), (100003, 7)])
def test_decode_random_ragged_sizes(engine, n, m):
    codes = synth.simulate_hardcalls(n, m, seed=n + m, missing_rate=0.03)
    geno, maf = engine.decode_2bit(synth.pack_2bit(codes), n)
    want = lut_decode(codes)
    assert bits_equal(geno, want)
    tot = np.nansum(want, 1)
    cnt = (~np.isnan(want)).sum(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        wm = tot / (2.0 * cnt)
    wm = np.where(wm > .5, 1 - wm, wm)
    assert bits_equal(maf, wm)
    rf = np.random.default_rng(n).random(n) < 0.25
    geno, _ = engine.decode_2bit(synth.pack_2bit(codes), n, row_filter=rf)
    assert bits_equal(geno, want[:, ~rf])


def make_pheno(n, seed, n_items=5):
    rng = np.random.default_rng(seed)
    f = rng.normal(size=n)
    items = {f"i{k}": (0.4 + 0.1 * k) * f + rng.normal(size=n) * (1 + 0.1 * k) + 0.3 * k for k in range(1, n_items + 1)}
    ph = pd.DataFrame(items)
    ph["sexm"] = rng.normal(size=n) + 1.0
    return ph


def oracle_rows(flat, ph, codes):
    import fit_oracle as fo
    cols = {c: ph[c].to_numpy(float) for c in ph.columns}
    geno = np.array([0.0, 1.0, 2.0, np.nan])[codes]
    return [fo.fit_snp_cumulants(flat, cols, g) for g in geno]


def compare(res, want, names):
    from gwsem_b200.engine import STATUS_TEXT
    for i, w in enumerate(want):
        if w["catch1"]:
            assert res.catch_code[i] != 0, (i, w["catch1"])
            continue
        assert res.catch_code[i] == 0, (i, res.catch_code[i])
        assert STATUS_TEXT[int(res.status[i])] == w["status"], (i, res.status[i], w["status"])
        assert res.n_used[i] == w["n"]
        se_w = np.sqrt(np.diag(w["vcov"]))
        se_g = np.sqrt(np.diag(res.vcov[i]))
        # estimates: 1e-4 relative, with the estimate's own standard error as the floor for
        # parameters whose true value is ~0 (a relative test on a noise-level number is meaningless)
        tol = EST_RTOL * np.maximum(np.abs(w["est"]), se_w)
        assert np.all(np.abs(res.est[i] - w["est"]) <= tol), (i, names, res.est[i], w["est"])
        assert np.allclose(se_g, se_w, rtol=SE_RTOL, atol=0), (i, se_g, se_w)
        assert np.allclose(res.vcov[i], w["vcov"], rtol=SE_RTOL, atol=SE_RTOL * np.outer(se_w, se_w).max())


@pytest.mark.parametrize("builder", ["item", "item2", "onefac", "onefacres", "twofac", "item_gxe"])
@pytest.mark.parametrize("n,missing", [(500, 0.0), (3001, 0.02)])
def test_fit_matches_oracle(engine, builder, n, missing):
    from gwsem_b200.engine import Model
    ph = make_pheno(n, seed=n, n_items=6)
    items = [f"i{k}" for k in range(1, 7)]
    model = {"item": lambda: buildItem(ph, "i1"),
             "item2": lambda: buildItem(ph, ["i1", "i2"]),
             "onefac": lambda: buildOneFac(ph[items[:5]], items[:5]),
             "onefacres": lambda: buildOneFacRes(ph[items[:5]], items[:5]),
             "twofac": lambda: buildTwoFac(ph[items], items[:3], items[3:]),
             "item_gxe": lambda: buildItem(ph, "i2", gxe=["sexm"])}[builder]()
    flat = flatten(model)
    m_snps = 24
    codes = synth.simulate_hardcalls(n, m_snps, seed=7 + n, missing_rate=missing)
    codes[3] = 0                      # monomorphic: minVariance catch
    codes[4, : n // 2] = 1; codes[4, n // 2:] = 0
    data = {c: model.data[c].to_numpy(float) for c in model.data.columns if c != "snp"}
    gm = Model(engine, flat, data)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    want = oracle_rows(flat, model.data, codes)
    compare(res, want, flat["par_names"])
    # PLINK 1 coding of the same genotypes gives the same answers
    bed = np.array([3, 2, 0, 1], np.uint8)[codes]
    res2 = gm.fit_2bit(synth.pack_2bit(bed), plink1=True)
    assert np.array_equal(res2.est.view(np.uint64), res.est.view(np.uint64))
    gm.close()


def test_fit_with_row_filter_and_missing_phenotypes(engine):
    from gwsem_b200.engine import Model
    n_src = 2000
    rng = np.random.default_rng(11)
    rf = rng.random(n_src) < 0.2
    ph = make_pheno(int((~rf).sum()), seed=5)
    ph.loc[ph.index[::37], "i2"] = np.nan        # naAction='omit' on phenotype rows
    model = buildOneFacRes(ph[["i1", "i2", "i3", "i4"]], ["i1", "i2", "i3", "i4"])
    flat = flatten(model)
    codes = synth.simulate_hardcalls(n_src, 16, seed=3, missing_rate=0.01)
    gm = Model(engine, flat, {c: model.data[c].to_numpy(float) for c in model.data.columns if c != "snp"})
    gm.set_row_filter(rf)
    res = gm.fit_2bit(synth.pack_2bit(codes))
    want = oracle_rows(flat, model.data, codes[:, ~rf])
    compare(res, want, flat["par_names"])
    gm.close()


@pytest.mark.parametrize("builder,n", [("item", 4097), ("onefacres", 3001), ("item", 100_000)])
def test_tcgen05_and_mma_sync_class_sums_are_bit_identical(engine, builder, n, monkeypatch):
    """The class sums are an exact integer contraction: the tcgen05 kernel (stats_tc5.cu) and the mma.sync kernel
    (stats_imma.cu) must give the same bits whatever the tile shape or summation order, at the bench's cohort size
    too, with missing calls and in PLINK 1 coding."""
    from gwsem_b200.engine import Model
    ph = make_pheno(n, seed=n, n_items=5)
    items = [f"i{k}" for k in range(1, 6)]
    model = buildItem(ph, "i1") if builder == "item" else buildOneFacRes(ph[items], items)
    flat = flatten(model)
    codes = synth.simulate_hardcalls(n, 40, seed=n + 1, missing_rate=0.01)
    codes[5, :] = np.where(codes[5, :] == 3, 0, codes[5, :])   # one SNP without a missing call
    data = {c: model.data[c].to_numpy(float) for c in model.data.columns if c != "snp"}
    packed = synth.pack_2bit(codes)
    bed = synth.pack_2bit(np.array([3, 2, 0, 1], np.uint8)[codes])
    out = {}
    for kernel in ("imma", "tc5"):
        monkeypatch.setenv("GWSEM_CS_KERNEL", kernel)
        gm = Model(engine, flat, data)
        out[kernel] = (gm.fit_2bit(packed), gm.fit_2bit(bed, plink1=True))
        gm.close()
    for a, b in zip(out["imma"], out["tc5"]):
        for f in ("est", "vcov", "fit", "maf"):
            assert np.array_equal(getattr(a, f).view(np.uint64), getattr(b, f).view(np.uint64)), f
        for f in ("n_used", "status", "catch_code"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert np.array_equal(out["imma"][0].est.view(np.uint64), out["imma"][1].est.view(np.uint64))
