"""GPU parity of the file-level path through the C ABI: genotype sources (every PGEN record
type, PLINK 1 .bed, BGEN layouts 1 and 2) decoded on the device, and the whole GWAS() loop with
its checkpoint log -- modelled on the reference's tests/testthat/test-formats.R,
test-rowFilter.R, test-model.R and test-gxe.R."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import EXTDATA, GOLDEN
from gwsem_b200 import GWAS, buildItem, buildOneFac, buildOneFacRes, flatten, loadResults, signif, signifGxE, isSuspicious, synth
from gwsem_b200._lib import GwsemError
from gwsem_b200.engine import Source
from test_oracle_decode import _synthetic_all_types

pytestmark = pytest.mark.gpu


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


FIXTURES = [("example.pgen", "pgen", -1), ("example.bed", "pgen", 500), ("test2.pgen", "pgen", -1),
            ("group1.bed", "pgen", 100), ("group2.bed", "pgen", 100)]


@pytest.mark.parametrize("name,method,n", FIXTURES)
def test_fixture_columns_bit_exact(engine, name, method, n):
    """What loadRow hands to OpenMx, for every SNP of every bundled file == the reference's own
    pgenlib output (golden), with and without rowFilter; MAF bit-exact too."""
    gold = np.load(os.path.join(GOLDEN, name.replace(".", "_") + ".npz"))
    with Source(os.path.join(EXTDATA, name), method, n) as s:
        assert s.num_variants == gold["geno"].shape[0] and s.load_counter == 1
        geno, maf = s.load_rows(engine, np.arange(s.num_variants))
        assert bits_equal(geno, gold["geno"]) and bits_equal(maf, gold["maf"])
        geno, maf = s.load_rows(engine, np.arange(s.num_variants), gold["row_filter"])
        assert bits_equal(geno, gold["geno_filtered"]) and bits_equal(maf, gold["maf_filtered"])
        idx = [5, 0, 5, s.num_variants - 1]                     # arbitrary order, repeats
        geno, _ = s.load_rows(engine, idx)
        assert bits_equal(geno, gold["geno"][idx])


def test_random_row_filters_subset_exactly(engine):
    """tests/testthat/test-rowFilter.R:24-32: 25 random filters, column == full[!toSkip]."""
    rng = np.random.default_rng(2)
    for name, method, n in (("example.pgen", "pgen", -1), ("example.bgen", "bgen", 500), ("test2.pgen", "pgen", -1)):
        with Source(os.path.join(EXTDATA, name), method, n) as s:
            full, _ = s.load_rows(engine, [0])
            for _ in range(25):
                skip = rng.random(s.num_samples) < .2
                sub, _ = s.load_rows(engine, [0], skip)
                assert bits_equal(sub[0], full[0][~skip])


def test_bgen_fixture(engine):
    gold = np.load(os.path.join(GOLDEN, "example_bgen.npz"))
    with Source(os.path.join(EXTDATA, "example.bgen"), "bgen", 500) as s:
        assert (s.num_variants, s.num_samples) == (199, 500)
        assert s.checkpoint_columns == ["SNP", "RSID", "CHR", "BP", "A2", "A1", "MAF"]
        geno, maf = s.load_rows(engine, np.arange(199))
        assert bits_equal(geno, gold["geno"])
        assert np.allclose(maf, gold["maf"], rtol=1e-13, atol=0)   # layout 1 dosages are dyadic: sums exact
        assert [s.context(i)[1] for i in range(199)] == list(gold["rsid"])   # .bgi order, not file order
        geno, _ = s.load_rows(engine, np.arange(199), gold["row_filter"])
        assert bits_equal(geno, gold["geno_filtered"])
    with pytest.raises(GwsemError, match="has 500 samples"):
        Source(os.path.join(EXTDATA, "example.bgen"), "bgen", 400)


@pytest.mark.parametrize("n", [97, 1000, 3001])
def test_every_pgen_record_type_on_device(engine, tmp_path, n):
    from oracle_lib import OrcPgen
    rng = np.random.default_rng(n)
    m = 48
    codes, kinds, dos, dk, hp = _synthetic_all_types(n, m, rng)
    p = str(tmp_path / "v.pgen")
    synth.write_pgen_variable(p, codes, kinds, dos, dk, hp, seed=n)
    o = OrcPgen(p)
    rf = rng.random(n) < .3
    with Source(p, "pgen", -1) as s:
        for filt in (None, rf):
            want, wmaf = o.load_rows(range(m), filt)
            got, gmaf = s.load_rows(engine, np.arange(m), filt)
            assert bits_equal(got, want) and bits_equal(gmaf, wmaf)
        order = rng.permutation(m)[:20]                       # LD records without their base in the batch
        got, _ = s.load_rows(engine, order)
        assert bits_equal(got, o.load_rows(order)[0])
    for fname, writer, hint in (("f.pgen", lambda q: synth.write_pgen_fixed(q, codes), -1),
                                ("f.bed", lambda q: synth.write_bed(q, codes), n),
                                ("d.pgen", lambda q: synth.write_pgen_fixed_dosage(q, codes, dos), -1)):
        q = str(tmp_path / fname)
        writer(q)
        o2 = OrcPgen(q, hint)
        with Source(q, "pgen", hint) as s:
            got, gmaf = s.load_rows(engine, np.arange(m), rf)
            want, wmaf = o2.load_rows(range(m), rf)
            assert bits_equal(got, want) and bits_equal(gmaf, wmaf)


@pytest.mark.parametrize("layout,bits,compression", [(1, 16, 1), (1, 16, 0), (2, 8, 1), (2, 16, 0), (2, 11, 1), (2, 32, 1)])
def test_bgen_layouts_on_device(engine, tmp_path, layout, bits, compression):
    from oracle_lib import OrcBgen
    rng = np.random.default_rng(bits + layout)
    m, n = 9, 1203
    denom = 32768 if layout == 1 else (1 << bits) - 1
    a = rng.integers(0, denom + 1, (m, n))
    b = (rng.random((m, n)) * (denom - a)).astype(np.int64)
    a[:, ::17] = 0
    b[:, ::17] = 0
    p = str(tmp_path / "t.bgen")
    synth.write_bgen(p, np.stack([a, b], -1), layout=layout, bits=bits, compression=compression,
                     positions=list(rng.permutation(m) * 10 + 5))
    o = OrcBgen(p)
    rf = rng.random(n) < .25
    with Source(p, "bgen", n) as s:
        for filt in (None, rf):
            want, wmaf = o.load_rows(range(m), filt)
            got, gmaf = s.load_rows(engine, np.arange(m), filt)
            assert bits_equal(got, want)
            assert np.allclose(gmaf, wmaf, rtol=1e-12, atol=0)   # parallel vs sequential double sum
        assert [s.context(i)[3] for i in range(m)] == [str(o.context(i)["BP"]) for i in range(m)]


def test_bgen_errors_match_reference(engine, tmp_path):
    probs = np.full((2, 40, 2), 60, np.int64)
    miss = np.zeros((2, 40), bool)
    miss[1, 3] = True
    p = str(tmp_path / "m.bgen")
    synth.write_bgen(p, probs, layout=2, bits=8, missing=miss)
    with Source(p, "bgen", 40) as s:
        s.load_rows(engine, [0])
        with pytest.raises(GwsemError, match="not implemented"):          # loader.cpp:52-55
            s.load_rows(engine, [1])
        rf = np.zeros(40, np.int32)
        rf[3] = 1
        s.load_rows(engine, [1], rf)
    synth.write_bgen(p, probs, layout=2, bits=8, phased=True)
    with Source(p, "bgen", 40) as s, pytest.raises(GwsemError, match="set_min_max_ploidy"):   # loader.cpp:18-26
        s.load_rows(engine, [0])


def example_pheno():
    ph = pd.read_csv(os.path.join(EXTDATA, "example.psam"), sep="\t")
    ph = ph.rename(columns={"#FID": "FID"})
    rng = np.random.default_rng(1)
    for k in range(1, 6):
        ph[f"i{k}"] = ph["phenotype"] * (0.4 + 0.1 * k) + rng.normal(size=len(ph))
    return ph


def check_log_against_oracle(log, flat, pheno_cols, geno, snps):
    import fit_oracle as fo
    names = flat["par_names"]
    for row, snp in zip(log.itertuples(index=False), snps):
        w = fo.fit_snp_cumulants(flat, pheno_cols, geno[snp])
        r = row._asdict() if hasattr(row, "_asdict") else dict(zip(log.columns, row))
        rec = dict(zip(log.columns, row))
        assert rec["MxComputeLoop1"] == snp + 1
        if w["catch1"]:
            assert "observed variance less" in rec["catch1"] or "not positive definite" in rec["catch1"]
            continue
        assert rec["catch1"] == "" and rec["statusCode"] == w["status"]
        se = np.sqrt(np.diag(w["vcov"]))
        for k, nme in enumerate(names):
            assert abs(rec[nme] - w["est"][k]) <= 1e-4 * max(abs(w["est"][k]), se[k])
            assert rec[f"V{nme}:{nme}"] == pytest.approx(w["vcov"][k, k], rel=2e-3)


def read_log(path):
    d = pd.read_csv(path, sep="\t", quoting=3, dtype={"catch1": str, "statusCode": str})
    d["catch1"] = d["catch1"].fillna("")
    return d


def test_gwas_example_pgen_dosages(engine, tmp_path):
    """C1-style run on the bundled example.pgen (dosage records): every row of the log against the
    oracle fed with the reference-decoded column; log layout as loadResults expects."""
    ph = example_pheno()
    model = buildItem(ph, "i3")
    out = str(tmp_path / "out.log")
    res = GWAS(model, os.path.join(EXTDATA, "example.pgen"), out)
    gold = np.load(os.path.join(GOLDEN, "example_pgen.npz"))
    log = read_log(out)
    assert len(log) == 199 and res.loadCounter == 1                       # test-formats.R:47-48
    assert bits_equal(res.last_snp, gold["geno"][198])
    assert list(log["SNP"][:3]) == ["RSID_2", "RSID_3", "RSID_4"] and set(log["A1"]) == {"A"} and set(log["A2"]) == {"G"}
    assert np.allclose(log["MAF"], np.round(gold["maf"], 8), atol=1e-12)
    flat = flatten(model)
    check_log_against_oracle(log, flat, {"i3": ph["i3"].to_numpy()}, gold["geno"], range(199))
    lr = loadResults(out, "snp_to_i3")
    assert list(lr.columns) == ["MxComputeLoop1", "CHR", "BP", "SNP", "A1", "A2", "statusCode", "catch1",
                                "snp_to_i3", "Vsnp_to_i3:snp_to_i3"]       # test-formats.R:82-84
    lr = signif(lr, "snp_to_i3")
    assert np.isfinite(lr["P"][~isSuspicious(lr)]).all()


def test_gwas_snp_selection_startfrom_and_errors(engine, tmp_path):
    ph = example_pheno()
    model = buildOneFacRes(ph[[f"i{k}" for k in range(1, 6)]], [f"i{k}" for k in range(1, 6)])
    out = str(tmp_path / "out.log")
    pgen = os.path.join(EXTDATA, "example.pgen")
    GWAS(model, pgen, out, SNP=[3, 4, 5])
    log = read_log(out)
    assert list(log["SNP"]) == ["RSID_4", "RSID_5", "RSID_6"]             # test-model.R:120
    assert list(log["MxComputeLoop1"]) == [3, 4, 5]
    GWAS(model, pgen, out, startFrom=190)
    assert len(read_log(out)) == 10                                        # test-formats.R:89-94
    with pytest.raises(GwsemError, match=r"out of data \(record 200 requested but only 199 in file\)"):
        GWAS(model, pgen, out, SNP=[200])                                  # test-model.R:60-66
    with pytest.raises(ValueError, match="does not match the number of rows of observed"):
        GWAS(model, pgen, out, rowFilter={"OneFacRes": np.arange(495) % 5 == 0}, SNP=[1])   # test-rowFilter.R:14-22
    with pytest.raises(ValueError, match="Rejected are any values"):
        GWAS(model, pgen, out, "oops")


def test_gwas_bed_and_bgen_agree_with_oracle(engine, tmp_path):
    ph = example_pheno()
    items = [f"i{k}" for k in range(1, 5)]
    model = buildOneFac(ph[items], items)
    flat = flatten(model)
    cols = {c: ph[c].to_numpy() for c in items}
    out = str(tmp_path / "out.log")
    GWAS(model, os.path.join(EXTDATA, "example.bed"), out)
    bed = read_log(out)
    gold = np.load(os.path.join(GOLDEN, "example_bed.npz"))
    assert len(bed) == 199 and list(bed["A1"][:2]) == ["A", "A"]
    check_log_against_oracle(bed, flat, cols, gold["geno"], range(199))
    GWAS(model, os.path.join(EXTDATA, "example.bgen"), out)
    bg = read_log(out)
    gb = np.load(os.path.join(GOLDEN, "example_bgen.npz"))
    assert len(bg) == 199 and list(bg["RSID"]) == list(gb["rsid"])
    check_log_against_oracle(bg, flat, cols, gb["geno"], range(199))
    pg_maf = np.load(os.path.join(GOLDEN, "example_pgen.npz"))["maf"]
    assert abs(np.abs(pg_maf - bed["MAF"]).sum() - 1.447) < 0.03           # test-formats.R:110


def test_gwas_row_filter_and_gxe(engine, tmp_path):
    ph = example_pheno()
    ph["mod"] = np.random.default_rng(5).normal(size=len(ph)) + 1
    skip = np.random.default_rng(7).random(len(ph)) < .2
    sub = ph[~skip].reset_index(drop=True)
    model = buildItem(sub, "i3", gxe=["mod"])
    out = str(tmp_path / "out.log")
    res = GWAS(model, os.path.join(EXTDATA, "example.pgen"), out, rowFilter={"OneItem": skip}, SNP=list(range(1, 41)))
    gold = np.load(os.path.join(GOLDEN, "example_pgen.npz"))
    assert bits_equal(res.last_snp, gold["geno"][39][~skip])                # test-rowFilter.R:31
    log = read_log(out)
    flat = flatten(model)
    check_log_against_oracle(log, flat, {"i3": sub["i3"].to_numpy(), "mod": sub["mod"].to_numpy()},
                             gold["geno"][:, ~skip], range(40))
    assert "Vsnp_mod_to_i3:snp_to_i3" in log.columns or "Vsnp_to_i3:snp_mod_to_i3" in log.columns   # test-gxe.R:67-77
    lr = loadResults(out, "snp_mod_to_i3")
    g = signifGxE(lr, "snp_mod_to_i3", level=.5)
    assert np.isfinite(g["ME.SE"][~isSuspicious(g)]).all() and (~isSuspicious(g)).sum() > 30


def test_multigroup_independent_groups(engine, tmp_path):
    """tests/testthat/test-multigroup.R: two buildItem models with their own parameter labels and
    their own .bed file, fitted jointly as a multigroup model, reproduce the two separate fits."""
    import os
    import pandas as pd
    from conftest import EXTDATA
    from gwsem_b200 import GWAS, MultigroupModel, buildItem, coefNames, mxRename, omxSetParameters
    logs = {}
    groups = []
    for k in (1, 2):
        fam = pd.read_csv(os.path.join(EXTDATA, f"group{k}.fam"), sep=r"\s+", header=None,
                          names=["V1", "V2", "V3", "V4", "V5", "V6"])
        g = mxRename(buildItem(fam, depVar="V6"), f"sample{k}")
        g = omxSetParameters(g, labels=coefNames(g), newlabels=[f"S{k}{n}" for n in coefNames(g)])
        groups.append(g)
        out = str(tmp_path / f"g{k}.log")
        GWAS(g, os.path.join(EXTDATA, f"group{k}.bed"), out, SNP=[1, 2, 3])
        logs[k] = pd.read_csv(out, sep="\t", quoting=3)
    mg = MultigroupModel("mg", *groups)
    out = str(tmp_path / "mg.log")
    GWAS(mg, {"sample1": os.path.join(EXTDATA, "group1.bed"), "sample2": os.path.join(EXTDATA, "group2.bed")}, out,
         SNP=[1, 2, 3])
    both = pd.read_csv(out, sep="\t", quoting=3)
    assert len(both) == 3 and list(both["MxComputeLoop1"]) == [1, 2, 3]
    for k in (1, 2):
        for name in coefNames(groups[k - 1]):
            assert np.allclose(both[name], logs[k][name], rtol=1e-12, equal_nan=True)
            assert np.allclose(both[f"V{name}:{name}"], logs[k][f"V{name}:{name}"], rtol=1e-12, equal_nan=True)
    assert np.allclose(both["fit"], logs[1]["fit"] + logs[2]["fit"])
    # ... and an independent *joint* fit: mxFitFunctionMultigroup minimises the sum of the groups' WLS fit functions
    # over the concatenated parameter vector; a generic optimiser on that sum (no knowledge that it separates) must
    # land on the multigroup log's estimates (tests/testthat/test-multigroup.R:38, tolerance 1e-4)
    import fit_oracle as fo
    from scipy import optimize
    from gwsem_b200.gwas import model_columns
    from gwsem_b200.model import flatten
    parts = []
    for k in (1, 2):
        flat = flatten(groups[k - 1])
        data, _ = model_columns(groups[k - 1])
        geno = np.load(os.path.join(os.path.dirname(EXTDATA), f"group{k}_bed.npz"))["geno"]
        parts.append((flat, data, geno, fo.FlatRAM(flat)))
    for row, snp in zip(both.to_dict("records"), (0, 1, 2)):
        pieces = []
        for flat, data, geno, ram in parts:
            X = fo.snp_columns(flat, data, geno[snp])
            sv, gamma, n = fo.cumulants_stats(X)
            pieces.append((ram, sv, n * np.linalg.inv(gamma), len(flat["par_names"])))

        def joint(theta):
            total, at = 0.0, 0
            for ram, sv, W, P in pieces:
                r = sv - fo.sigma_cumulants(ram, theta[at:at + P])
                total += r @ W @ r
                at += P
            return total

        x0 = np.concatenate([np.asarray(flat["par_start"], float) for flat, *_ in parts])
        best = optimize.minimize(joint, x0, method="Nelder-Mead", options=dict(xatol=1e-10, fatol=1e-14, maxiter=20000, maxfev=20000))
        names = [n for flat, *_ in parts for n in flat["par_names"]]
        assert best.fun == pytest.approx(row["fit"], rel=1e-6, abs=1e-9)
        for nme, v in zip(names, best.x):
            assert row[nme] == pytest.approx(v, rel=1e-4, abs=1e-6), (snp, nme)
    shared = MultigroupModel("mg2", groups[0], mxRename(groups[0], "again"))
    with pytest.raises(NotImplementedError, match="shared"):
        GWAS(shared, {"sample1": os.path.join(EXTDATA, "group1.bed"), "again": os.path.join(EXTDATA, "group1.bed")}, out, SNP=[1])


@pytest.mark.parametrize("fmt", ["pgen", "bed", "bgen"])
def test_gwas_log_does_not_depend_on_the_batch_size(engine, tmp_path, fmt):
    """The run loop keeps three batches in flight (one on the device, one being copied out, one being written) and only
    brings the logged covariance entries back: the log of a run cut into 13 batches of 16 SNPs equals the log of the
    same run in one batch, byte for byte once the time stamp column is removed, for every kind of source (read-ahead of
    fixed-width records, per-record staging of dosage records, bgen blocks)."""
    ph = example_pheno()
    model = buildOneFac(ph, [f"i{k}" for k in range(1, 6)])
    path = {"pgen": os.path.join(EXTDATA, "example.pgen"), "bed": os.path.join(EXTDATA, "example.bed"),
            "bgen": os.path.join(EXTDATA, "example.bgen")}[fmt]
    logs = []
    for k, batch in enumerate((0, 16, 7)):
        out = str(tmp_path / f"out{k}.log")
        GWAS(model, path, out, batch_snps=batch)
        with open(out) as f:
            rows = [line.rstrip("\n").split("\t") for line in f]
        ts = rows[0].index("timestamp")
        logs.append([[x for i, x in enumerate(r) if i != ts] for r in rows])
    assert len(logs[0]) == 200
    assert logs[1] == logs[0] and logs[2] == logs[0]
