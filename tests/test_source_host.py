"""Host-side behaviour of the genotype sources (no GPU needed): error texts the reference's providers raise."""
import os
import struct

import pytest

from conftest import EXTDATA
from gwsem_b200 import synth
from gwsem_b200._lib import GwsemError
from gwsem_b200.engine import Source
from oracle_lib import REF_SO, RefPgen


def multiallelic_pgen(path, n=20, m=3):
    """A variable-width .pgen whose header control byte says "ALT-allele counts stored, 1 byte each" (bits 4-5 = 1):
    what plink 2 writes as soon as one variant has more than two alleles."""
    rec = bytes((n + 3) // 4)
    ctrl = 0x10 | 0x00          # 4-bit vrtypes + 1-byte record lengths, alt-allele counts stored, nonref flags unstored
    with open(path, "wb") as f:
        f.write(b"\x6c\x1b\x10" + struct.pack("<II", m, n) + bytes([ctrl]))
        first = 12 + 8 + (m + 1) // 2 + m + m
        f.write(struct.pack("<Q", first))
        f.write(bytes((m + 1) // 2))               # vrtypes: all 0
        f.write(bytes([len(rec)] * m))             # record lengths
        f.write(bytes([1, 2, 1]))                  # ALT allele counts: the second variant has three alleles
        f.write(rec * m)


def test_multiallelic_pgen_is_rejected_like_the_reference(tmp_path):
    """src/loader.cpp:323-327 opens every file with allele_idx_offsets == NULL; pgenlib refuses files that store
    ALT-allele counts (src/pgenlib_read.cc:1172-1178), so aux1 tracks (SkipAux1, 6605-6620) are never reached."""
    p = str(tmp_path / "multi.pgen")
    multiallelic_pgen(p)
    with pytest.raises(GwsemError, match="allele_idx_offsets must be allocated before PgfiInitPhase2"):
        Source(p, "pgen", -1)
    if os.path.exists(REF_SO):
        with pytest.raises(RuntimeError, match="allele_idx_offsets must be allocated before PgfiInitPhase2"):
            RefPgen(p)


def test_counts_without_a_gpu():
    with Source(os.path.join(EXTDATA, "example.pgen"), "pgen", -1) as s:
        assert s.num_variants == 199 and s.num_samples == 500
    with Source(os.path.join(EXTDATA, "example.bgen"), "bgen", -1) as s:
        assert s.num_variants == 199
    with Source(os.path.join(EXTDATA, "group1.bed"), "pgen", -1) as s:
        assert s.num_variants is None          # tests/testthat/test-formats.R:25-33: NA without the sample count
