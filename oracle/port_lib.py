"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/marginals_port.c, the compiled (-O3) restatement of
oracle/marginals_oracle.py::fit_snp_marginals.  Used by tests/test_oracle_marginals_port.py (port == numpy oracle) and by
bench.py's CPU-baseline legs; never by gwsem_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(ORACLE_DIR, "libgwsem_port.so")
MAXV, MAXK, MAXP = 12, 12, 96
STATUS = {0: "OK", 1: "iteration limit/blue", 2: "nonzero gradient/red", 3: "not convex/red", -1: "NA"}


class PortModel(C.Structure):
    _fields_ = [("n", C.c_int), ("p", C.c_int), ("K", C.c_int),
                ("col", C.c_void_p * MAXV), ("kind", C.c_int * MAXV), ("n_cat", C.c_int * MAXV),
                ("x", C.c_void_p * MAXK), ("exo_free", (C.c_int * MAXK) * MAXV),
                ("nv", C.c_int), ("nm", C.c_int), ("n_cells", C.c_int), ("P", C.c_int),
                ("cells", C.c_void_p), ("par_start", C.c_void_p), ("par_lbound", C.c_void_p),
                ("n_thr", C.c_int), ("thr", C.c_void_p), ("exo_latent", C.c_int * MAXK),
                ("min_variance", C.c_double), ("snp_is_exo", C.c_int)]


class PortResult(C.Structure):
    _fields_ = [("est", C.c_double * MAXP), ("vcov", C.c_double * (MAXP * MAXP)), ("fit", C.c_double),
                ("n", C.c_int), ("iterations", C.c_int), ("status", C.c_int), ("catch_code", C.c_int), ("catch_var", C.c_int)]


_lib = None


def port_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PORT_SO):
            subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port"], check=True)
        _lib = C.CDLL(PORT_SO)
        _lib.gwsem_port_sizeof_model.restype = C.c_size_t
        _lib.gwsem_port_sizeof_result.restype = C.c_size_t
        assert _lib.gwsem_port_sizeof_model() == C.sizeof(PortModel), "PortModel layout"
        assert _lib.gwsem_port_sizeof_result() == C.sizeof(PortResult), "PortResult layout"
        _lib.gwsem_port_fit_marginals.argtypes = [C.POINTER(PortModel), C.POINTER(PortResult)]
    return _lib


class MarginalsPort:
    """One model (as gwsem_b200.model.flatten describes it) + its phenotype columns; fit(snp) = one row of the GWAS loop."""

    def __init__(self, flat, pheno_cols, n_cats):
        names = flat["manifest"]
        if any(m.startswith("snp_") for m in list(names) + list(flat["exo_pred"])):
            raise NotImplementedError("gxe product columns are not part of the compiled port")
        self.flat, self.names = flat, names
        self.P = len(flat["par_names"])
        m = PortModel()
        self._keep = []
        first = next(iter(pheno_cols.values()))
        self.n = len(first)
        m.n, m.p, m.K = self.n, len(names), len(flat["exo_pred"])
        self.snp_slot = names.index("snp") if "snp" in names else -1
        for j, nm in enumerate(names):
            m.kind[j] = 1 if n_cats.get(nm, 0) > 1 else 0
            m.n_cat[j] = int(n_cats.get(nm, 0))
            if nm != "snp":
                a = np.ascontiguousarray(pheno_cols[nm], dtype=np.float64)
                self._keep.append(a)
                m.col[j] = a.ctypes.data
        self.snp_exo = -1
        for k, c in enumerate(flat["exo_pred"]):
            if c == "snp":
                self.snp_exo = k
            else:
                a = np.ascontiguousarray(pheno_cols[c], dtype=np.float64)
                self._keep.append(a)
                m.x[k] = a.ctypes.data
            m.exo_latent[k] = len(names) + flat["latent"].index(c)
        ef = flat["exo_free"] if flat["exo_free"] is not None else np.zeros((len(names), 0), bool)
        for j in range(len(names)):
            for k in range(m.K):
                m.exo_free[j][k] = int(bool(ef[j][k]))
        self._cells = np.ascontiguousarray(np.asarray(flat["cells"], dtype=np.float64).reshape(-1, 5))
        self._start = np.ascontiguousarray(flat["par_start"], dtype=np.float64)
        self._lb = np.ascontiguousarray(flat["par_lbound"], dtype=np.float64)
        self._thr = np.ascontiguousarray(np.asarray(flat["thresholds"], dtype=np.float64).reshape(-1, 4))
        m.nv, m.nm, m.n_cells, m.P = flat["nv"], flat["nm"], len(self._cells), self.P
        m.cells, m.par_start, m.par_lbound = self._cells.ctypes.data, self._start.ctypes.data, self._lb.ctypes.data
        m.n_thr, m.thr = len(self._thr), self._thr.ctypes.data
        m.min_variance = float(flat["min_variance"])
        m.snp_is_exo = self.snp_exo
        self.m = m
        self.res = PortResult()

    def fit(self, snp):
        snp = np.ascontiguousarray(snp, dtype=np.float64)
        if self.snp_slot >= 0:
            self.m.col[self.snp_slot] = snp.ctypes.data
        if self.snp_exo >= 0:
            self.m.x[self.snp_exo] = snp.ctypes.data
        rc = port_lib().gwsem_port_fit_marginals(C.byref(self.m), C.byref(self.res))
        if rc != 0:
            raise RuntimeError("model too large for the compiled port")
        P, r = self.P, self.res
        catch = ""
        if r.catch_code == 1:
            var = "snp" if r.catch_var == -2 else self.names[r.catch_var]
            catch = f"{self.flat['name']}.data: '{var}' has observed variance less than {self.flat['min_variance']:g}"
        elif r.catch_code == 2:
            catch = f"{self.flat['name']}.data: asymptotic covariance matrix is not positive definite"
        return dict(est=np.array(r.est[:P]), vcov=np.array(r.vcov[:P * P]).reshape(P, P), status=STATUS[r.status], catch1=catch,
                    fit=r.fit, n=r.n, iterations=r.iterations)
