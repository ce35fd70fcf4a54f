"""ctypes bindings for the test oracles (TEST INFRASTRUCTURE, never imported by gwsem_b200/;
only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it).

* ``RefPgen``    -- oracle/_ref/libgwsem_ref.so: the reference's own pgenlib, driven as
                    src/loader.cpp:290-406 drives it.
* ``OrcPgen`` / ``OrcBgen`` -- oracle/libgwsem_oracle.so: our plain-C restatement.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(ORACLE_DIR)
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libgwsem_ref.so")
ORC_SO = os.path.join(ORACLE_DIR, "libgwsem_oracle.so")
NA_BITS = 0x7FF00000000007A2


def build_oracles():
    """Compile oracle/ (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


def _load(path):
    if not os.path.exists(path):
        build_oracles()
    return C.CDLL(path)


_ref = None
_orc = None


def ref_lib():
    global _ref
    if _ref is None:
        L = _load(REF_SO)
        L.ref_pgen_open.restype = C.c_void_p
        L.ref_pgen_open.argtypes = [C.c_char_p, C.c_int]
        L.ref_last_error.restype = C.c_char_p
        L.ref_pgen_variant_ct.argtypes = [C.c_void_p]
        L.ref_pgen_sample_ct.argtypes = [C.c_void_p]
        L.ref_pgen_load_row.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_pgen_get1d.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_pgen_close.argtypes = [C.c_void_p]
        _ref = L
    return _ref


def orc_lib():
    global _orc
    if _orc is None:
        L = _load(ORC_SO)
        L.orc_last_error.restype = C.c_char_p
        L.orc_pgen_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        for f in ("orc_pgen_variant_ct", "orc_pgen_sample_ct", "orc_pgen_close",
                  "orc_bgen_variant_ct", "orc_bgen_sample_ct", "orc_bgen_layout", "orc_bgen_close"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_pgen_vrtype.argtypes = [C.c_void_p, C.c_int]
        L.orc_pgen_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_pgen_load_row.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_bgen_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.orc_bgen_load_row.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_bgen_file_index.argtypes = [C.c_void_p, C.c_int]
        L.orc_bgen_field.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_bgen_field.restype = C.c_char_p
        L.orc_bgen_position.argtypes = [C.c_void_p, C.c_int]
        L.orc_bgen_position.restype = C.c_uint32
        L.orc_bgen_file_offset.argtypes = [C.c_void_p, C.c_int]
        L.orc_bgen_file_offset.restype = C.c_uint64
        _orc = L
    return _orc


def _filter_arg(row_filter, n):
    if row_filter is None:
        return None, n
    rf = np.ascontiguousarray(np.asarray(row_filter).astype(np.int32))
    assert rf.shape == (n,)
    return rf, int((rf == 0).sum())


class _Rows:
    def load_rows(self, snps, row_filter=None):
        rf, dest = _filter_arg(row_filter, self.N)
        out = np.empty((len(snps), dest))
        maf = np.empty(len(snps))
        for k, i in enumerate(snps):
            out[k], maf[k] = self.load_row(int(i), rf)
        return out, maf


class RefPgen(_Rows):
    def __init__(self, path, sample_ct=-1):
        self.L = ref_lib()
        self.h = self.L.ref_pgen_open(os.fsencode(path), int(sample_ct))
        if not self.h:
            raise RuntimeError(self.L.ref_last_error().decode())
        self.M = self.L.ref_pgen_variant_ct(self.h)
        self.N = self.L.ref_pgen_sample_ct(self.h)

    def load_row(self, index, rf=None):
        dest = self.N if rf is None else int((rf == 0).sum())
        out = np.empty(dest)
        maf = C.c_double()
        r = self.L.ref_pgen_load_row(self.h, index, None if rf is None else rf.ctypes.data,
                                     out.ctypes.data, C.byref(maf))
        if r < 0:
            raise RuntimeError(self.L.ref_last_error().decode())
        return out, maf.value

    def get1d(self, index):
        """(codes[N] uint8, dosage[N] uint16 with 0xFFFF = none)"""
        n = self.N
        geno = np.zeros((n + 3) // 4, np.uint8)
        present = np.zeros((n + 7) // 8, np.uint8)
        dmain = np.zeros(n, np.uint16)
        ct = self.L.ref_pgen_get1d(self.h, index, geno.ctypes.data, present.ctypes.data, dmain.ctypes.data)
        if ct < 0:
            raise RuntimeError(self.L.ref_last_error().decode())
        codes = (np.repeat(geno, 4) >> np.tile(np.arange(4, dtype=np.uint8) * 2, len(geno))) & 3
        codes = codes[:n].astype(np.uint8)
        dosage = np.full(n, 0xFFFF, np.uint16)
        if ct:
            bits = np.unpackbits(present, bitorder="little")[:n].astype(bool)
            dosage[bits] = dmain[:ct]
        return codes, dosage

    def close(self):
        if self.h:
            self.L.ref_pgen_close(self.h)
            self.h = None

    __del__ = close


class OrcPgen(_Rows):
    def __init__(self, path, sample_ct=-1):
        self.L = orc_lib()
        h = C.c_void_p()
        if self.L.orc_pgen_open(os.fsencode(path), int(sample_ct), C.byref(h)):
            raise RuntimeError(self.L.orc_last_error().decode())
        self.h = h
        self.M = self.L.orc_pgen_variant_ct(h)
        self.N = self.L.orc_pgen_sample_ct(h)

    def vrtype(self, i):
        return self.L.orc_pgen_vrtype(self.h, i)

    def load_row(self, index, rf=None):
        dest = self.N if rf is None else int((rf == 0).sum())
        out = np.empty(dest)
        maf = C.c_double()
        r = self.L.orc_pgen_load_row(self.h, index, None if rf is None else rf.ctypes.data,
                                     out.ctypes.data, C.byref(maf))
        if r < 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return out, maf.value

    def get(self, index):
        codes = np.zeros(self.N, np.uint8)
        dosage = np.zeros(self.N, np.uint16)
        if self.L.orc_pgen_get(self.h, index, codes.ctypes.data, dosage.ctypes.data):
            raise RuntimeError(self.L.orc_last_error().decode())
        return codes, dosage

    def close(self):
        if self.h:
            self.L.orc_pgen_close(self.h)
            self.h = None

    __del__ = close


class OrcBgen(_Rows):
    def __init__(self, path):
        self.L = orc_lib()
        h = C.c_void_p()
        if self.L.orc_bgen_open(os.fsencode(path), C.byref(h)):
            raise RuntimeError(self.L.orc_last_error().decode())
        self.h = h
        self.M = self.L.orc_bgen_variant_ct(h)
        self.N = self.L.orc_bgen_sample_ct(h)
        self.layout = self.L.orc_bgen_layout(h)

    def load_row(self, index, rf=None):
        dest = self.N if rf is None else int((rf == 0).sum())
        out = np.empty(dest)
        maf = C.c_double()
        r = self.L.orc_bgen_load_row(self.h, index, None if rf is None else rf.ctypes.data,
                                     out.ctypes.data, C.byref(maf))
        if r < 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return out, maf.value

    def context(self, index):
        f = lambda w: self.L.orc_bgen_field(self.h, index, w).decode()
        return dict(SNP=f(0), RSID=f(1), CHR=f(2), BP=int(self.L.orc_bgen_position(self.h, index)),
                    A1=f(3), A2=f(4), file_index=self.L.orc_bgen_file_index(self.h, index),
                    file_offset=int(self.L.orc_bgen_file_offset(self.h, index)))

    def close(self):
        if self.h:
            self.L.orc_bgen_close(self.h)
            self.h = None

    __del__ = close


def bits_equal(a, b):
    """Bit-exact comparison of two float64 arrays (NA payloads included)."""
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))
