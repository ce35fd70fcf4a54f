"""GWAS(): the reference's driver API (R/model.R:24-45, 122-315) over the native engine.

``GWAS(model, snpData, out, SNP=..., startFrom=..., rowFilter=...)`` keeps the argument
names, defaults and error messages of the reference; instead of assembling an OpenMx compute
plan it hands the flattened model, the genotype source and the job description to
``gwsem_gwas_run`` (C ABI), which runs the whole per-SNP loop on the GPU and writes the same
tab separated checkpoint log that ``loadResults`` reads.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import check, lib
from .engine import Engine, Model, Source
from .model import flatten

_engines = {}


def _engine(device):
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]


def processSnpData(snpData):
    """R/model.R:24-45: file extension -> data provider method."""
    single = isinstance(snpData, str)
    items = [snpData] if single else list(snpData)
    ext, stem, method = [], [], []
    for s in items:
        pieces = s.split(".")
        if len(pieces) < 2:
            raise ValueError(f"Please rename snpData '{s}' to the form file.ext where ext reflects the format of the data")
        e = pieces[-1]
        if e in ("pgen", "bed"):
            m = "pgen"
        elif e == "bgen":
            m = "bgen"
        else:
            raise ValueError(f"Unrecognized file extension '{e}' inferred from snpData '{s}'")
        ext.append(e)
        stem.append(".".join(pieces[:-1]))
        method.append(m)
    return dict(snpFileExt=ext, stem=stem, method=method)


def numAvailableRecords(snpData):
    """R/model.R:210-224.  NA (None) for a .bed, whose variant count needs the sample count."""
    items = [snpData] if isinstance(snpData, str) else list(snpData)
    psd = processSnpData(items)
    out = {}
    for path, method in zip(items, psd["method"]):
        with Source(path, method, -1) as s:
            out[path] = s.num_variants
    return out


def buildAnalysesPlan(snpData, sliceSize):
    """R/model.R:249-274: split each file's SNP range into jobs of about sliceSize SNPs."""
    import pandas as pd
    items = [snpData] if isinstance(snpData, str) else list(snpData)
    recs = numAvailableRecords(items)
    rows = []
    for path in items:
        n = recs[path]
        slices = max(n // int(round(sliceSize)), 1)
        if slices == 1:
            rows.append(dict(path=path, begin=1, end=n, count=n, slice=1))
            continue
        breaks = np.floor(np.linspace(1, n, slices + 1)).astype(int)
        begin = breaks[:-1]
        end = np.append(begin[1:] - 1, n)
        for k in range(slices):
            rows.append(dict(path=path, begin=int(begin[k]), end=int(end[k]), count=int(end[k] - begin[k] + 1), slice=k + 1))
    return pd.DataFrame(rows)


def model_columns(model):
    """Observed columns as float64 (ordinal factors -> 0-based codes, NA -> NaN) and their level counts."""
    import pandas as pd
    data, n_categories = {}, {}
    for c in model.data.columns:
        col = model.data[c]
        if c == "snp" or c.startswith("snp_"):
            continue
        if isinstance(col.dtype, pd.CategoricalDtype):
            codes = col.cat.codes.to_numpy().astype(np.float64)
            codes[codes < 0] = np.nan
            data[c] = codes
            n_categories[c] = len(col.cat.categories)
        elif col.dtype.kind in "fiub":
            data[c] = col.to_numpy(dtype=np.float64)
    return data, n_categories


def _offdiag(model, names):
    """R/model.R:176-185: off-diagonal vcov entries needed for gxe follow-up (signifGxE)."""
    if not model.gxe:
        return []
    gxe = ["snp_" + e for e in model.gxe]
    col = model.vars.index("snp")
    outcome = [model.vars[r] for r in range(len(model.vars)) if model.A["free"][r, col]]
    want = [f"snp_to_{o}" for o in outcome] + [f"{g}_to_{o}" for o in outcome for g in gxe + list(model.gxe)]
    return [n for n in want if n in names]


def _gwas_multigroup(model, snpData, out, SNP, startFrom, rowFilter, device, batch_snps):
    """GWAS(mg, c(sample1=file1, sample2=file2)) (R/model.R:132-163, tests/testthat/test-multigroup.R).
    With no parameter label shared between the groups the multigroup fit (a sum of the groups' fit
    functions) separates: every group is fitted on its own file and the rows are joined by SNP index.
    Shared labels would couple the groups' iterations and are not implemented."""
    import tempfile
    import pandas as pd
    from .model import MultigroupModel
    if not isinstance(snpData, dict) or not snpData:
        raise ValueError("Must provide model names for snpData. For example,\n {'" + str(getattr(model, "name", "model")) + "': path}")
    if not isinstance(model, MultigroupModel):
        raise ValueError("several snpData files need a multigroup model (MultigroupModel(name, g1, g2, ...))")
    by_name = {m.name: m for m in model.submodels}
    for k in snpData:
        if k not in by_name:
            raise ValueError(f"Cannot find model associated with snpData: '{k}'")
    if isinstance(rowFilter, dict):
        unknown = [k for k in rowFilter if k not in by_name]
        if unknown:
            raise ValueError("Cannot find model associated with rowFilter: " + ", ".join(f"'{k}'" for k in unknown))
    elif rowFilter is not None:
        raise ValueError("rowFilter must be a dict keyed by sub-model name for a multigroup model")
    seen = {}
    for m in model.submodels:
        for lab in flatten(m)["par_names"]:
            if lab in seen:
                raise NotImplementedError(f"parameter '{lab}' is shared by sub-models '{seen[lab]}' and '{m.name}': "
                                          "groups coupled by shared parameters are not implemented")
            seen[lab] = m.name
    parts, results = [], []
    with tempfile.TemporaryDirectory() as tmp:
        for name, path in snpData.items():
            sub_out = os.path.join(tmp, name + ".log")
            results.append(GWAS(by_name[name], path, sub_out, SNP=SNP, startFrom=startFrom,
                                rowFilter=(rowFilter or {}).get(name), device=device, batch_snps=batch_snps))
            parts.append(pd.read_csv(sub_out, sep="\t", quoting=3, keep_default_na=False,
                                     dtype=str))
    n = min(len(p_) for p_ in parts)
    first = parts[0].iloc[:n].copy()
    shared = ["OpenMxEvals", "iterations", "timestamp", "MxComputeLoop1", "fit", "fitUnits", "sampleSize", "statusCode",
              "CHR", "BP", "SNP", "A1", "A2", "RSID", "MAF", "catch1"]
    merged = first[[c for c in ["OpenMxEvals", "iterations", "timestamp", "MxComputeLoop1"] if c in first]].copy()
    for p_ in parts:
        for c in p_.columns:
            if c not in shared and not c.startswith("V"):
                merged[c] = p_[c].iloc[:n].to_numpy()
    num = lambda col: pd.to_numeric(pd.Series(col).replace("NA", np.nan), errors="coerce")
    merged["fit"] = sum(num(p_["fit"].iloc[:n].to_numpy()) for p_ in parts).map(lambda v: "NA" if pd.isna(v) else repr(float(v)))
    merged["fitUnits"] = first["fitUnits"].to_numpy()
    merged["sampleSize"] = sum(pd.to_numeric(p_["sampleSize"].iloc[:n]) for p_ in parts).astype(int).astype(str)
    status = first["statusCode"].to_numpy().copy()
    catch1 = first["catch1"].to_numpy().copy()
    for p_ in parts[1:]:
        st, ca = p_["statusCode"].iloc[:n].to_numpy(), p_["catch1"].iloc[:n].to_numpy()
        status = np.where(status == "OK", st, status)
        catch1 = np.where(catch1 == "", ca, catch1)
    merged["statusCode"] = status
    for p_ in parts:
        for c in p_.columns:
            if c.startswith("V") and ":" in c:
                merged[c] = p_[c].iloc[:n].to_numpy()
    for c in ["CHR", "BP", "SNP", "A1", "A2", "RSID", "MAF"]:   # variant context and MAF of the first group's file
        if c in first:
            merged[c] = first[c].to_numpy()
    merged["catch1"] = catch1
    merged.to_csv(out, sep="\t", index=False, quoting=3)
    print(f"Done. See '{out}' for results")
    return GWASResult(model, out, n, None, sum(r.loadCounter for r in results))


def call_arguments(model, snpData, out="out.log", SNP=None, startFrom=1, rowFilter=None, device=0, batch_snps=0):
    """The six arguments of .Call("gwsem_run", spec, data, path, method, rowFilter, job) (src/gwsem_call.c) as the
    reference-side glue R/flatten.R assembles them from an MxModel -- flattenModel(), manifestColumns(), jobList().
    Returned as plain Python values; dump_call() writes them to a text file that the C harness of `make -C src check`
    turns back into SEXPs, so the .Call layer is exercised end to end without R."""
    psd = processSnpData(snpData)
    method, ext, stem = psd["method"][0], psd["snpFileExt"][0], psd["stem"][0]
    flat = flatten(model)
    data, n_categories = model_columns(model)
    nm = flat["nm"]
    n_rows = len(model.data)
    kinds, cols = [], []
    for name in flat["manifest"]:
        if name == "snp":
            kinds.append(1); cols.append(None)
        elif name.startswith("snp_") and name[4:] in flat["gxe"]:
            kinds.append(2); cols.append(data[name[4:]])
        else:
            kinds.append(0); cols.append(data[name])
    exo, exo_kind = [], []
    for name in flat["exo_pred"]:
        if name == "snp" and "snp" not in data:
            exo.append(None); exo_kind.append(1)
        elif name.startswith("snp_") and name[4:] in flat["gxe"] and name not in data:
            exo.append(data[name[4:]]); exo_kind.append(2)
        else:
            exo.append(data[name]); exo_kind.append(0)
    exo_free = (flat["exo_free"].astype(np.uint8) if flat["exo_free"] is not None else np.zeros((nm, 0), np.uint8))
    fit = {("WLS", "cumulants"): 0, ("WLS", "marginals"): 1}.get((flat["fitfun"], flat["continuous_type"]), 2)
    spec = dict(n_vars=flat["nv"], n_manifest=nm, n_params=len(flat["par_names"]), n_cells=len(flat["cells"]),
                cells=np.asarray(flat["cells"], float).reshape(-1), par_start=np.asarray(flat["par_start"], float),
                par_lbound=np.asarray(flat["par_lbound"], float), fit=fit, min_variance=float(flat["min_variance"]),
                n_rows=n_rows, manifest_kind=np.asarray(kinds, np.int32),
                n_categories=np.asarray([n_categories.get(m, 0) for m in flat["manifest"]], np.int32),
                exo=exo, exo_latent=np.asarray([nm + flat["latent"].index(c) for c in flat["exo_pred"]], np.int32),
                exo_kind=np.asarray(exo_kind, np.int32), exo_free=exo_free.reshape(-1),
                n_thresholds=len(flat["thresholds"]), thresholds=np.asarray(flat["thresholds"], float).reshape(-1))
    names = flat["par_names"]
    off = _offdiag(model, names)
    pairs = [(names.index(a), names.index(b)) for i, a in enumerate(off) for b in off[i + 1:]]
    job = dict(device=int(device), snp=None if SNP is None else np.asarray(SNP, np.int32).reshape(-1),
               start_from=int(startFrom), out=out, par_names=list(names),
               context_path={"pgen": stem + ".pvar", "bed": stem + ".bim"}.get(ext),
               vcov_pairs=np.asarray(pairs, np.int32).reshape(-1) if pairs else None, batch_snps=int(batch_snps),
               model_name=model.name, manifest_names=list(flat["manifest"]))
    rf = None if rowFilter is None else np.asarray(rowFilter).astype(bool)
    return dict(spec=spec, data=cols, path=snpData, method=method, rowFilter=rf, job=job)


def dump_call(args, path):
    """One value per line: `<list> <name> <type> <n> v...`; doubles as C99 hex floats (exact), strings %-quoted."""
    from urllib.parse import quote

    def line(lst, name, v):
        if v is None:
            return f"{lst} {name} null 0"
        if isinstance(v, str):
            return f"{lst} {name} str 1 {quote(v, safe='/')}"
        if isinstance(v, (list, tuple)) and all(isinstance(x, str) for x in v):
            return f"{lst} {name} str {len(v)} " + " ".join(quote(x, safe='/') for x in v)
        if isinstance(v, (bool, np.bool_)):
            return f"{lst} {name} lgl 1 {int(v)}"
        if isinstance(v, (int, np.integer)):
            return f"{lst} {name} int 1 {int(v)}"
        if isinstance(v, float):
            return f"{lst} {name} real 1 {float(v).hex()}"
        a = np.asarray(v)
        if a.dtype == bool:
            return f"{lst} {name} lgl {a.size} " + " ".join(str(int(x)) for x in a.reshape(-1))
        if a.dtype == np.uint8:
            return f"{lst} {name} raw {a.size} " + " ".join(str(int(x)) for x in a.reshape(-1))
        if a.dtype.kind in "iu":
            return f"{lst} {name} int {a.size} " + " ".join(str(int(x)) for x in a.reshape(-1))
        return f"{lst} {name} real {a.size} " + " ".join("nan" if x != x else float(x).hex() for x in a.reshape(-1))

    out = []
    for k, v in args["spec"].items():
        if k == "exo":
            out.append(f"spec n_exo int 1 {len(v)}")
            out += [line("exo", str(j), c) for j, c in enumerate(v)]
        else:
            out.append(line("spec", k, v))
    out += [line("data", str(j), c) for j, c in enumerate(args["data"])]
    out += [line("top", "path", args["path"]), line("top", "method", args["method"]), line("top", "rowFilter", args["rowFilter"])]
    out += [line("job", k, v) for k, v in args["job"].items()]
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


class GWASResult:
    """What the reference's GWAS() returns invisibly: the model with the last SNP loaded."""

    def __init__(self, model, out, rows, last_snp, load_counter):
        self.model, self.out, self.rows_written = model, out, rows
        self.last_snp = last_snp          # m$data$observed$snp after the run
        self.loadCounter = load_counter   # m$compute$steps[[2]]$debug$loadCounter


class ComputePlan:
    """The per-SNP loop prepareComputePlan attaches to a model (R/model.R:122-195), as data: the engine runs
    exactly these steps for every SNP -- reset starts, load the SNP column, load its context, fit, standard
    errors, one checkpoint row -- in batches on the device (gwsem_gwas_run)."""

    def __init__(self, model, snpData, out, SNP, startFrom, rowFilter):
        self.snpData, self.out, self.SNP, self.startFrom, self.rowFilter = snpData, out, SNP, startFrom, rowFilter
        ml = flatten(model)["fitfun"] != "WLS" if isinstance(snpData, str) else None
        self.steps = ["ST: mxComputeSetOriginalStarts", f"LD: mxComputeLoadData({snpData!r}, column='snp')",
                      "LC: mxComputeLoadContext(CHR, BP, SNP, A1, A2)",
                      "TC: mxComputeTryCatch(GD: mxComputeGradientDescent" +
                      (", ND: mxComputeNumericDeriv, SE: mxComputeStandardError, HQ: mxComputeHessianQuality)" if ml
                       else ", SE: mxComputeStandardError)"),
                      f"CK: mxComputeCheckpoint({out!r}, standardErrors=FALSE, vcov=TRUE)"]

    def __repr__(self):
        return "mxComputeLoop[\n  " + "\n  ".join(self.steps) + "\n]"


def prepareComputePlan(model, snpData, out="out.log", *dots, SNP=None, startFrom=1, rowFilter=None):
    """R/model.R:122-195: returns the model with its compute plan (model.compute); GWAS() builds the same plan
    and runs it.  mxRun(model) of the reference = run_plan(model) here."""
    if dots:
        raise ValueError("Rejected are any values passed in the '...' argument")
    if isinstance(snpData, str):
        processSnpData(snpData)   # rejects unknown formats as the reference does
    model.compute = ComputePlan(model, snpData, out, SNP, startFrom, rowFilter)
    return model


def run_plan(model, device=0, batch_snps=0):
    """mxRun() of a model prepared by prepareComputePlan."""
    plan = getattr(model, "compute", None)
    if plan is None:
        raise ValueError("model has no compute plan: call prepareComputePlan first")
    return GWAS(model, plan.snpData, plan.out, SNP=plan.SNP, startFrom=plan.startFrom, rowFilter=plan.rowFilter,
                device=device, batch_snps=batch_snps)


def GWAS(model, snpData, out="out.log", *dots, SNP=None, startFrom=1, rowFilter=None, device=0, batch_snps=0, _timing=None):
    """R/model.R:305-315.  (_timing: optional dict that receives the seconds spent inside gwsem_gwas_run.)"""
    if dots:
        raise ValueError("Rejected are any values passed in the '...' argument")
    if not isinstance(snpData, str):
        return _gwas_multigroup(model, snpData, out, SNP=SNP, startFrom=startFrom, rowFilter=rowFilter, device=device,
                                batch_snps=batch_snps)
    psd = processSnpData(snpData)
    method, ext, stem = psd["method"][0], psd["snpFileExt"][0], psd["stem"][0]
    if isinstance(rowFilter, dict):
        unknown = [k for k in rowFilter if k != model.name]
        if unknown:
            raise ValueError("Cannot find model associated with rowFilter: " + ", ".join(f"'{k}'" for k in unknown))
        rowFilter = rowFilter.get(model.name)
    n_rows = len(model.data)
    if rowFilter is not None:
        rowFilter = np.asarray(rowFilter).astype(bool)
        if int((~rowFilter).sum()) != n_rows:
            raise ValueError(f"rowFilter keeps {int((~rowFilter).sum())} rows, which does not match the number of rows "
                             f"of observed data ({n_rows})")
        src_rows = len(rowFilter)
    else:
        src_rows = n_rows
    flat = flatten(model)
    data, n_categories = model_columns(model)
    if flat["fitfun"] != "WLS":
        if any(n_categories.get(m, 0) > 0 for m in flat["manifest"]):
            raise NotImplementedError("fitfun='ML' with ordinal items is not implemented (continuous items only)")
    eng = _engine(device)
    with Source(snpData, method, src_rows) as src:
        gm = Model(eng, flat, data, n_categories)
        try:
            gm.set_row_filter(rowFilter, src_rows)
            names = flat["par_names"]
            off = _offdiag(model, names)
            pairs = [(names.index(a), names.index(b)) for i, a in enumerate(off) for b in off[i + 1:]]
            ctx = {"pgen": stem + ".pvar", "bed": stem + ".bim"}.get(ext)
            snp_arr = None if SNP is None else np.ascontiguousarray(np.asarray(SNP, dtype=np.int32).reshape(-1))
            c_names = (C.c_char_p * len(names))(*[n.encode() for n in names])
            c_manifest = (C.c_char_p * flat["nm"])(*[n.encode() for n in flat["manifest"]])
            c_pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1))
            job = _lib.Job(snp_arr.ctypes.data_as(C.POINTER(C.c_int32)) if snp_arr is not None else None,
                           -1 if snp_arr is None else len(snp_arr), int(startFrom), os.fsencode(out), c_names,
                           os.fsencode(ctx) if ctx else None,
                           c_pairs.ctypes.data_as(C.POINTER(C.c_int32)) if len(pairs) else None, len(pairs),
                           int(batch_snps), model.name.encode(), c_manifest)
            written = C.c_int64()
            import time
            t0 = time.perf_counter()
            check(lib().gwsem_gwas_run(gm.h, src.h, C.byref(job), C.byref(written)))
            if _timing is not None:
                _timing["gwas_run_s"] = time.perf_counter() - t0
            last = None
            if written.value:
                last_idx = int(snp_arr[-1]) - 1 if snp_arr is not None and len(snp_arr) else src.num_variants - 1
                last = src.load_rows(eng, [last_idx], rowFilter)[0][0]
            counter = src.load_counter
        finally:
            gm.close()
    print(f"Done. See '{out}' for results")
    return GWASResult(model, out, int(written.value), last, counter)
