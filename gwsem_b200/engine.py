"""Python handles over the C ABI: Engine (one per GPU), Model, result containers."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib

STATUS_TEXT = {0: "OK", 1: "OK/green", 4: "iteration limit/blue", 5: "not convex/red",
               6: "nonzero gradient/red", -1: "NA"}


def _ptr(a):
    """address of a numpy array, torch tensor (host or device) or None"""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


class Engine:
    def __init__(self, device=0):
        self.h = C.c_void_p()
        check(lib().gwsem_engine_create(int(device), C.byref(self.h)))
        self.device = device

    def set_stream(self, cuda_stream_ptr):
        check(lib().gwsem_engine_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    def synchronize(self):
        check(lib().gwsem_engine_synchronize(self.h))

    @property
    def launches(self):
        return int(lib().gwsem_engine_launch_count(self.h))

    @property
    def series_columns(self):
        """int8 digit columns of the series-path class-sum launches so far, summed over launches"""
        return int(lib().gwsem_engine_series_columns(self.h))

    KERNEL_FAMILIES = ("class_sums", "count_missing", "fit", "decode", "marg_stats", "marg_exact", "marg_series", "series_sums")

    def profile(self, enable=True):
        check(lib().gwsem_engine_profile(self.h, int(enable)))

    def profile_read(self):
        """{family: (launches, total_ms)} since the last read (CUDA events on the launch stream)"""
        n = (C.c_int64 * 8)()
        ms = (C.c_double * 8)()
        check(lib().gwsem_engine_profile_read(self.h, n, ms))
        return {k: (int(n[i]), float(ms[i])) for i, k in enumerate(self.KERNEL_FAMILIES)}

    def fp64_peak(self):
        """Measured DFMA throughput of this GPU right now, TFLOP/s."""
        v = C.c_double()
        check(lib().gwsem_engine_fp64_peak(self.h, C.byref(v)))
        return float(v.value)

    def decode_2bit(self, packed, n_samples, plink1=False, row_filter=None):
        """packed: uint8 [n_snps, pitch] host array -> (geno[n_snps, dest_rows], maf[n_snps])"""
        packed = np.ascontiguousarray(packed, np.uint8)
        n_snps, pitch = packed.shape
        rf = None if row_filter is None else np.ascontiguousarray(np.asarray(row_filter).astype(np.int32))
        dest = n_samples if rf is None else int((rf == 0).sum())
        out = np.empty((n_snps, dest))
        maf = np.empty(n_snps)
        check(lib().gwsem_decode_2bit(self.h, _ptr(packed), pitch, n_snps, n_samples, int(plink1), _ptr(rf),
                                      _ptr(out), _ptr(maf)))
        return out, maf

    def close(self):
        if self.h:
            lib().gwsem_engine_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FitResult:
    """Host-side structure-of-arrays result of a batch of SNPs."""

    def __init__(self, n_snps, n_params, alloc=np.empty):
        self.n, self.P = n_snps, n_params
        self.est = alloc((n_snps, n_params), dtype=np.float64)
        self.vcov = alloc((n_snps, n_params, n_params), dtype=np.float64)
        self.fit = alloc(n_snps, dtype=np.float64)
        self.maf = alloc(n_snps, dtype=np.float64)
        self.n_used = alloc(n_snps, dtype=np.int32)
        self.status = alloc(n_snps, dtype=np.int32)
        self.catch_code = alloc(n_snps, dtype=np.int32)
        self.catch_var = alloc(n_snps, dtype=np.int32)
        self.iterations = alloc(n_snps, dtype=np.int32)

    def c_struct(self):
        return _lib.Result(*[_ptr(getattr(self, f)) for f, _ in _lib.Result._fields_])


class Model:
    """A flattened RAM model (gwsem_b200.model.flatten) bound to an engine."""

    FIT_CODES = {("WLS", "cumulants"): 0, ("WLS", "marginals"): 1, ("ML", "marginals"): 2, ("ML", "cumulants"): 2}

    def __init__(self, engine, flat, data, n_categories=None):
        """data: {column name: float64 array over the model's rows (after rowFilter)}; ordinal items
        hold 0-based category codes (NaN = missing) and n_categories[name] their number of levels"""
        self.engine, self.flat = engine, flat
        self.names = list(flat["par_names"])
        self.P = len(self.names)
        nm = flat["nm"]
        self._keep = []
        kinds = np.zeros(nm, np.int32)
        ptrs = (C.POINTER(C.c_double) * nm)()
        n_rows = None
        for j, name in enumerate(flat["manifest"]):
            if name == "snp":
                kinds[j] = 1
                continue
            col = name
            if name.startswith("snp_") and name[4:] in flat["gxe"]:
                kinds[j] = 2
                col = name[4:]
            arr = np.ascontiguousarray(np.asarray(data[col], dtype=np.float64))
            self._keep.append(arr)
            n_rows = len(arr)
            ptrs[j] = arr.ctypes.data_as(C.POINTER(C.c_double))
        if n_rows is None:
            n_rows = len(next(iter(data.values())))
        # WLS marginals: ordinal codes, exogenous covariates, thresholds
        ncat = np.zeros(nm, np.int32)
        for j, name in enumerate(flat["manifest"]):
            ncat[j] = int((n_categories or {}).get(name, 0))
        exo = list(flat["exo_pred"])
        exo_ptrs = (C.POINTER(C.c_double) * max(len(exo), 1))()
        exo_kind = np.zeros(max(len(exo), 1), np.int32)
        for k, name in enumerate(exo):
            if name == "snp" and "snp" not in data:
                exo_kind[k] = 1
                continue  # NULL: the genotype column itself (buildItem with exogenous = TRUE)
            if name.startswith("snp_") and name[4:] in flat["gxe"] and name not in data:
                exo_kind[k] = 2   # genotype x moderator as a definition variable
                name = name[4:]
            arr = np.ascontiguousarray(np.asarray(data[name], dtype=np.float64))
            self._keep.append(arr)
            exo_ptrs[k] = arr.ctypes.data_as(C.POINTER(C.c_double))
        exo_latent = np.array([nm + flat["latent"].index(c) for c in exo], np.int32)
        exo_free = np.ascontiguousarray(flat["exo_free"].astype(np.uint8) if flat["exo_free"] is not None
                                        else np.zeros((nm, 0), np.uint8))
        thr = np.ascontiguousarray(flat["thresholds"], np.float64)
        self._keep += [ncat, exo_ptrs, exo_latent, exo_free, thr, exo_kind]
        cells = np.ascontiguousarray(flat["cells"], np.float64)
        start = np.ascontiguousarray(flat["par_start"], np.float64)
        lb = np.ascontiguousarray(flat["par_lbound"], np.float64)
        self._keep += [cells, start, lb, kinds, ptrs]
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        spec = _lib.ModelSpec(flat["nv"], nm, self.P, len(cells), dp(cells), dp(start), dp(lb),
                              self.FIT_CODES[(flat["fitfun"], flat["continuous_type"])],
                              float(flat["min_variance"]), n_rows,
                              kinds.ctypes.data_as(C.POINTER(C.c_int32)), ptrs,
                              ncat.ctypes.data_as(C.POINTER(C.c_int32)), len(exo), exo_ptrs,
                              exo_latent.ctypes.data_as(C.POINTER(C.c_int32)),
                              exo_free.ctypes.data_as(C.POINTER(C.c_uint8)), len(thr),
                              thr.ctypes.data_as(C.POINTER(C.c_double)),
                              exo_kind.ctypes.data_as(C.POINTER(C.c_int32)))
        self.n_rows = n_rows
        self.h = C.c_void_p()
        check(lib().gwsem_model_create(engine.h, C.byref(spec), C.byref(self.h)))

    def set_row_filter(self, row_filter, src_rows=None):
        rf = None if row_filter is None else np.ascontiguousarray(np.asarray(row_filter).astype(np.int32))
        n = len(rf) if rf is not None else int(src_rows)
        check(lib().gwsem_model_set_row_filter(self.h, _ptr(rf), n))

    def fit_2bit(self, packed, plink1=False, result=None):
        """Host path: packed uint8 [n_snps, pitch] (numpy or pinned torch) -> FitResult"""
        n_snps, pitch = packed.shape
        res = result or FitResult(n_snps, self.P)
        cs = res.c_struct()
        check(lib().gwsem_fit_2bit(self.h, _ptr(packed), pitch, n_snps, int(plink1), C.byref(cs)))
        return res

    def fit_2bit_device(self, d_packed_ptr, pitch, n_snps, d_result_struct, plink1=False):
        check(lib().gwsem_fit_2bit_device(self.h, C.c_void_p(d_packed_ptr), pitch, n_snps, int(plink1),
                                          C.byref(d_result_struct)))

    def fit_rows_device(self, d_geno_ptr, n_snps, d_result_struct):
        """Generic path on decoded doubles resident in HBM (dosages): d_geno[n_snps][n_pad] in file order, NaN =
        missing, n_pad = people rounded up to 32; the result's maf array must hold the MAFs on entry."""
        check(lib().gwsem_fit_rows_device(self.h, C.c_void_p(d_geno_ptr), n_snps, C.byref(d_result_struct)))

    def close(self):
        if self.h:
            lib().gwsem_model_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Source:
    """A genotype file (the reference's data providers, src/loader.cpp): pgen / bed / bgen."""

    def __init__(self, path, method=None, src_rows=-1):
        import os
        self.h = C.c_void_p()
        check(lib().gwsem_source_open(os.fsencode(path), method.encode() if method else None, int(src_rows),
                                      C.byref(self.h)))
        self.path = path

    @property
    def num_variants(self):
        n = lib().gwsem_source_num_variants(self.h)
        return None if n < 0 else n

    @property
    def num_samples(self):
        return lib().gwsem_source_num_samples(self.h)

    @property
    def load_counter(self):
        return lib().gwsem_source_load_counter(self.h)

    @property
    def checkpoint_columns(self):
        return lib().gwsem_source_checkpoint_columns(self.h).decode().split("\t")

    def context(self, index):
        buf = C.create_string_buffer(4096)
        check(lib().gwsem_source_context(self.h, int(index), buf, 4096))
        return buf.value.decode().split("\t") if buf.value else []

    def load_rows(self, engine, snp_index, row_filter=None):
        """loadRow for a batch of 0-based SNP indices -> (geno[n, dest_rows], maf[n])"""
        idx = np.ascontiguousarray(np.asarray(snp_index, dtype=np.int32))
        n = self.num_samples
        rf = None if row_filter is None else np.ascontiguousarray(np.asarray(row_filter).astype(np.int32))
        dest = n if rf is None else int((rf == 0).sum())
        out = np.empty((len(idx), dest))
        maf = np.empty(len(idx))
        check(lib().gwsem_source_load_rows(engine.h, self.h, _ptr(idx), len(idx), _ptr(rf), _ptr(out), _ptr(maf)))
        return out, maf

    def close(self):
        if self.h:
            lib().gwsem_source_close(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
