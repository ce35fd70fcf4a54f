"""Regenerates tests/golden/ from the reference (run in the build container only).

Inputs : the reference's bundled genotype fixtures, /root/reference/inst/extdata
         (copied verbatim into tests/golden/extdata/ -- they are data, not source).
Outputs: <name>.npz = what LoadDataPGENProvider2::loadRowImpl (reference
         src/loader.cpp:290-406) delivers for every SNP of the file, produced by
         oracle/_ref (the reference's own pgenlib compiled in place), i.e. by the
         reference itself:  geno[M, N] float64 (NA = R NA_real_ bit pattern),
         maf[M], and the same with a fixed row filter.
         example_bgen.npz = the .bgen fixture decoded by the C restatement
         (oracle/decode_oracle.c; the reference's bgen reader cannot be compiled here,
         SURVEY.md section 8c) -- pinned only through tests/testthat/test-formats.R:76,117.
"""
import hashlib
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
from oracle_lib import OrcBgen, RefPgen  # noqa: E402

SRC = "/root/reference/inst/extdata"
FILES = ["example.pgen", "example.pvar", "example.psam", "example.bed", "example.bim", "example.fam",
         "example.bgen", "example.bgen.bgi", "example.sample", "test2.pgen", "test2.pvar", "test2.psam",
         "group1.bed", "group1.bim", "group1.fam", "group2.bed", "group2.bim", "group2.fam"]
PGEN_LIKE = [("example.pgen", -1), ("example.bed", 500), ("test2.pgen", -1), ("group1.bed", 100),
             ("group2.bed", 100)]


def fixed_filter(n):
    return (np.random.default_rng(20261019).random(n) < 0.2).astype(np.int32)


def main():
    ext = os.path.join(HERE, "extdata")
    os.makedirs(ext, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(ext, f))
    sums = []
    for name, n in PGEN_LIKE:
        r = RefPgen(os.path.join(ext, name), n)
        geno, maf = r.load_rows(range(r.M))
        rf = fixed_filter(r.N)
        geno_f, maf_f = r.load_rows(range(r.M), rf)
        out = os.path.join(HERE, name.replace(".", "_") + ".npz")
        np.savez_compressed(out, geno=geno, maf=maf, row_filter=rf, geno_filtered=geno_f, maf_filtered=maf_f)
        sums.append((name, hashlib.sha256(geno.tobytes()).hexdigest()))
    b = OrcBgen(os.path.join(ext, "example.bgen"))
    geno, maf = b.load_rows(range(b.M))
    rf = fixed_filter(b.N)
    geno_f, maf_f = b.load_rows(range(b.M), rf)
    ctx = [b.context(i) for i in range(b.M)]
    np.savez_compressed(os.path.join(HERE, "example_bgen.npz"), geno=geno, maf=maf, row_filter=rf,
                        geno_filtered=geno_f, maf_filtered=maf_f,
                        rsid=np.array([c["RSID"] for c in ctx]), bp=np.array([c["BP"] for c in ctx]))
    sums.append(("example.bgen", hashlib.sha256(geno.tobytes()).hexdigest()))
    with open(os.path.join(HERE, "SHA256SUMS"), "w") as f:
        for name, s in sums:
            f.write(f"{s}  decoded:{name}\n")


if __name__ == "__main__":
    main()
