"""CPU tests of the decode oracle: the plain-C restatement (oracle/decode_oracle.c) against
(a) golden vectors produced by the reference's own pgenlib (tests/golden/*.npz, made by
tests/golden/make_golden.py from oracle/_ref), (b) oracle/_ref itself when it is built,
on the bundled fixtures and on synthetic files covering every record type, and (c) the numeric
pins of the reference's tests (tests/testthat/test-formats.R:110-117, 76)."""
import os
import sqlite3

import numpy as np
import pytest

from conftest import EXTDATA, GOLDEN
from gwsem_b200 import synth
from oracle_lib import REF_SO, OrcBgen, OrcPgen, RefPgen, bits_equal

HAVE_REF = os.path.exists(REF_SO) or os.path.isdir("/root/reference/src")
FIXTURES = [("example.pgen", -1), ("example.bed", 500), ("test2.pgen", -1), ("group1.bed", 100), ("group2.bed", 100)]


@pytest.mark.parametrize("name,n", FIXTURES)
def test_oracle_matches_reference_golden(name, n):
    gold = np.load(os.path.join(GOLDEN, name.replace(".", "_") + ".npz"))
    o = OrcPgen(os.path.join(EXTDATA, name), n)
    geno, maf = o.load_rows(range(o.M))
    assert bits_equal(geno, gold["geno"]) and bits_equal(maf, gold["maf"])
    geno, maf = o.load_rows(range(o.M), gold["row_filter"])
    assert bits_equal(geno, gold["geno_filtered"]) and bits_equal(maf, gold["maf_filtered"])


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("name,n", FIXTURES)
def test_reference_library_reproduces_golden(name, n):
    gold = np.load(os.path.join(GOLDEN, name.replace(".", "_") + ".npz"))
    r = RefPgen(os.path.join(EXTDATA, name), n)
    geno, maf = r.load_rows(range(r.M))
    assert bits_equal(geno, gold["geno"]) and bits_equal(maf, gold["maf"])


def test_reference_test_pins():
    """test-formats.R:110-111: sum|MAF_pgen - MAF_bed| == 1.447 +- 2%, max diff < .06;
    :117 bgen MAF within 1e-2 of pgen; :76 last SNP column of bgen == pgen's within 5e-5 (mean rel);
    test-formats.R:25-33 record counts; test-model.R:120 first variant is RSID_2."""
    pg = np.load(os.path.join(GOLDEN, "example_pgen.npz"))
    bed = np.load(os.path.join(GOLDEN, "example_bed.npz"))
    assert abs(np.abs(pg["maf"] - bed["maf"]).sum() - 1.447) < 0.02 * 1.447
    assert np.abs(pg["maf"] - bed["maf"]).max() < 0.06
    assert np.isnan(pg["geno"]).sum() == 2 and np.isnan(bed["geno"]).sum() == 4351
    b = OrcBgen(os.path.join(EXTDATA, "example.bgen"))
    assert (b.M, b.N, b.layout) == (199, 500, 1)
    geno, maf = b.load_rows(range(b.M))
    ids = [l.split("\t")[2] for l in open(os.path.join(EXTDATA, "example.pvar")).read().splitlines()[1:]]
    assert ids[0] == "RSID_2" and len(ids) == 199
    pos = {r: i for i, r in enumerate(ids)}
    perm = np.array([pos[b.context(i)["RSID"]] for i in range(b.M)])
    assert np.abs(maf - pg["maf"][perm]).max() < 1e-2
    last_b, last_p = geno[b.M - 1], pg["geno"][198]
    assert b.context(b.M - 1)["RSID"] == ids[198]
    ok = ~np.isnan(last_p) & ~np.isnan(last_b)
    assert np.abs(last_b[ok] - last_p[ok]).mean() / np.abs(last_p[ok]).mean() < 5e-5


def test_bgen_index_order_matches_bgi():
    """SNP i of a bgen file is row i of the .bgi Variant table in the reference's ORDER BY
    (src/bgen.cpp:649-663), not file order."""
    b = OrcBgen(os.path.join(EXTDATA, "example.bgen"))
    con = sqlite3.connect(os.path.join(EXTDATA, "example.bgen.bgi"))
    rows = con.execute("SELECT rsid, file_start_position FROM Variant ORDER BY chromosome, position, rsid, "
                       "allele1, allele2").fetchall()
    assert [(b.context(i)["RSID"], b.context(i)["file_offset"]) for i in range(b.M)] == rows
    assert b.context(0)["RSID"] == "RSID_101"


def _synthetic_all_types(n, m, rng):
    codes = np.zeros((m, n), np.uint8)
    kinds = []
    for i in range(m):
        k = [0, 1, 4, 6, 7, 5, 2, 3, 0, 1, 2, 3][i % 12] if i else 0
        if k == 0:
            c = synth.simulate_hardcalls(n, 1, 100 + i, missing_rate=0.02)[0]
        elif k == 1:
            c = synth.simulate_hardcalls(n, 1, 100 + i, maf_range=(0.1, 0.2))[0]
            cnt = np.bincount(c, minlength=4)
            order = np.argsort(-cnt, kind="stable")
            rare = np.nonzero((c == order[2]) | (c == order[3]))[0]
            c[rare[n // 8:]] = order[0]
        elif k in (4, 6, 7):
            c = np.full(n, k & 3, np.uint8)
            ids = rng.choice(n, rng.integers(0, n // 8 + 1), replace=False)
            c[ids] = rng.integers(0, 4, len(ids))
        elif k == 5:
            c = np.zeros(n, np.uint8)
        else:
            j = i - 1
            while kinds[j] in (2, 3):
                j -= 1
            b = codes[j].copy()
            ids = rng.choice(n, rng.integers(0, n // 8 + 1), replace=False)
            b[ids] = rng.integers(0, 4, len(ids))
            c = b if k == 2 else synth._invert(b)
        codes[i] = c
        kinds.append(k)
    dk = [[0, 0x20, 0x40, 0x60][rng.integers(0, 4)] for _ in range(m)]
    hp = [int(rng.integers(0, 3)) for _ in range(m)]
    dos = np.full((m, n), 0xFFFF, np.uint16)
    for i in range(m):
        if dk[i] == 0x20:
            ids = rng.choice(n, rng.integers(0, n // 8 + 1), replace=False)
            dos[i, ids] = rng.integers(0, 32769, len(ids))
        elif dk[i] == 0x40:
            dos[i] = rng.integers(0, 32769, n)
            dos[i, rng.random(n) < 0.05] = 0xFFFF
        elif dk[i] == 0x60:
            ids = rng.random(n) < 0.5
            dos[i, ids] = rng.integers(0, 32769, ids.sum())
    return codes, kinds, dos, dk, hp


@pytest.mark.parametrize("n", [97, 1000, 3001])
def test_every_pgen_record_type(tmp_path, n):
    rng = np.random.default_rng(n)
    m = 48
    codes, kinds, dos, dk, hp = _synthetic_all_types(n, m, rng)
    p = str(tmp_path / "v.pgen")
    vt = synth.write_pgen_variable(p, codes, kinds, dos, dk, hp, seed=n)
    assert {v & 7 for v in vt} == {0, 1, 2, 3, 4, 5, 6, 7}
    o = OrcPgen(p)
    for i in range(m):
        oc, od = o.get(i)
        assert np.array_equal(oc, codes[i]) and np.array_equal(od, dos[i]), (i, hex(vt[i]))
    rf = rng.random(n) < .3
    if HAVE_REF:
        r = RefPgen(p)
        for i in range(m):
            rc, rd = r.get1d(i)
            assert np.array_equal(rc, codes[i]) and np.array_equal(rd, dos[i]), (i, hex(vt[i]))
        for filt in (None, rf):
            a, ma = r.load_rows(range(m), filt)
            b, mb = o.load_rows(range(m), filt)
            assert bits_equal(a, b) and bits_equal(ma, mb)
    # the same hardcalls through the fixed-width, dosage and bed containers
    for path, writer, hint in ((tmp_path / "f.pgen", lambda q: synth.write_pgen_fixed(q, codes), -1),
                               (tmp_path / "f.bed", lambda q: synth.write_bed(q, codes), n),
                               (tmp_path / "d.pgen", lambda q: synth.write_pgen_fixed_dosage(q, codes, dos), -1)):
        writer(str(path))
        o2 = OrcPgen(str(path), hint)
        assert np.array_equal(o2.get(7)[0], codes[7])
        if HAVE_REF:
            r2 = RefPgen(str(path), hint)
            a, ma = r2.load_rows(range(m), rf)
            b, mb = o2.load_rows(range(m), rf)
            assert bits_equal(a, b) and bits_equal(ma, mb)


@pytest.mark.parametrize("layout,bits,compression", [(1, 16, 1), (1, 16, 0), (2, 8, 1), (2, 16, 0), (2, 11, 1), (2, 32, 1)])
def test_bgen_layouts(tmp_path, layout, bits, compression):
    rng = np.random.default_rng(bits + layout)
    m, n = 7, 203
    denom = 32768 if layout == 1 else (1 << bits) - 1
    a = rng.integers(0, denom + 1, (m, n))
    b = (rng.random((m, n)) * (denom - a)).astype(np.int64)
    a[:, ::17] = 0
    b[:, ::17] = 0                       # exact zero dosage -> NA (loader.cpp:42-43)
    probs = np.stack([a, b], -1)
    p = str(tmp_path / "t.bgen")
    pos = list(rng.permutation(m) * 10 + 5)
    synth.write_bgen(p, probs, layout=layout, bits=bits, compression=compression, positions=pos)
    o = OrcBgen(p)
    geno, maf = o.load_rows(range(m))
    order = np.argsort(pos, kind="stable")
    d = float(denom)
    want = (b[order] / d) * 1 + (a[order] / d) * 2
    want[want == 0] = np.nan
    assert np.array_equal(np.isnan(geno), np.isnan(want))
    assert np.array_equal(np.nan_to_num(geno), np.nan_to_num(want))
    assert [o.context(i)["BP"] for i in range(m)] == sorted(pos)
    rf = rng.random(n) < .3
    g2, _ = o.load_rows(range(m), rf)
    assert bits_equal(g2, geno[:, ~rf])


def test_bgen_missing_and_phased_are_errors(tmp_path):
    probs = np.full((2, 10, 2), 60, np.int64)
    miss = np.zeros((2, 10), bool)
    miss[1, 3] = True
    p = str(tmp_path / "m.bgen")
    synth.write_bgen(p, probs, layout=2, bits=8, missing=miss)
    o = OrcBgen(p)
    o.load_row(0)
    with pytest.raises(RuntimeError, match="not implemented"):   # loader.cpp:52-55
        o.load_row(1)
    rf = np.zeros(10, np.int32)
    rf[3] = 1
    o.load_row(1, rf)                                            # skipped sample: no throw
    synth.write_bgen(p, probs, layout=2, bits=8, phased=True)
    with pytest.raises(RuntimeError, match="set_min_max_ploidy"):  # loader.cpp:18-26
        OrcBgen(p).load_row(0)


def test_out_of_data_error(tmp_path):
    """loader.cpp:369-372, tests/testthat/test-model.R:60-66"""
    o = OrcPgen(os.path.join(EXTDATA, "example.pgen"))
    with pytest.raises(RuntimeError, match=r"out of data \(record 200 requested but only 199 in file\)"):
        o.load_row(199)
