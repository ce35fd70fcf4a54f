"""Model builders: the reference's R API (R/model.R) mirrored in Python.

``buildItem`` / ``buildOneFac`` / ``buildOneFacRes`` / ``buildTwoFac`` return a
:class:`RAMModel` holding exactly what the reference's builders put into their
OpenMx ``MxModel`` -- manifest / latent variables, the RAM ``A`` / ``S`` / ``M``
matrices with their free / fixed pattern, labels, start values and lower bounds,
thresholds, the exogenous-covariate ``exoFree`` pattern, the gxe product columns
and the fit function -- because that object *is* the specification the per-SNP
engine has to honour.  No OpenMx here: the engine
behind ``GWAS()`` consumes the flattened model (``RAMModel.flatten``).

Argument names, defaults, validation messages and parameter labels follow the
reference (file:line cited per function) so the reference's tests read the same.
"""
from dataclasses import dataclass, field

import numpy as np
import pandas as pd

defaultExogenous = True  # R/model.R:1


def calcMinVar(minMAF):  # R/model.R:10
    return 2 * minMAF * (1 - minMAF)


def _is_factor(col):
    return isinstance(col.dtype, pd.CategoricalDtype)


def mxNormalQuantiles(nBreaks, mean=0.0, sd=1.0):
    """OpenMx::mxNormalQuantiles: qnorm(seq(0,1,length.out=nBreaks+2))[2:(nBreaks+1)]."""
    from scipy.stats import norm
    p = np.linspace(0, 1, nBreaks + 2)[1:-1]
    return norm.ppf(p, loc=mean, scale=sd)


@dataclass
class Path:
    """One ``mxPath`` call (from, to, arrows, free, values, labels, lbound, connect)."""
    frm: list
    to: list = None
    arrows: int = 1
    free: object = True
    values: object = None
    labels: object = None
    lbound: object = None
    connect: str = "single"


def mxPath(frm, to=None, arrows=1, free=True, values=None, labels=None, lbound=None, connect="single"):
    aslist = lambda x: None if x is None else (list(x) if isinstance(x, (list, tuple, np.ndarray, pd.Index)) else [x])
    return Path(aslist(frm), aslist(to), arrows, free, values, labels, lbound, connect)


@dataclass
class RAMModel:
    name: str
    manifestVars: list
    latentVars: list
    data: pd.DataFrame                    # mxData(observed=..., type='raw')
    fitfun: str = "WLS"                   # "WLS" | "ML"
    continuousType: str = "marginals"     # mxFitFunctionWLS(allContinuousMethod=...)
    minVariance: float = 0.0
    naAction: str = "omit"
    gxe: list = field(default_factory=list)        # moderators E: data column snp_E = snp * E
    hasMeans: bool = True
    exoFree: pd.DataFrame = None                   # manifest x covariate logical
    exoPred: list = field(default_factory=list)    # exogenous covariates (definition variables)
    thresholds: dict = field(default_factory=dict) # item -> dict(values, labels, free)

    def __post_init__(self):
        v = self.vars
        n = len(v)
        z = lambda fill, dt: np.full((n, n), fill, dt)
        self.A = dict(values=z(0.0, float), free=z(False, bool), labels=z(None, object), lbound=z(np.nan, float))
        self.S = dict(values=z(0.0, float), free=z(False, bool), labels=z(None, object), lbound=z(np.nan, float))
        m = lambda fill, dt: np.full((1, n), fill, dt)
        self.M = dict(values=m(0.0, float), free=m(False, bool), labels=m(None, object), lbound=m(np.nan, float))

    @property
    def vars(self):
        return list(self.manifestVars) + list(self.latentVars)

    # -- mxModel(type='RAM', paths) -------------------------------------------------
    def add_paths(self, paths):
        idx = {v: i for i, v in enumerate(self.vars)}
        for p in paths:
            frm = p.frm
            to = p.to if p.to is not None else (p.frm if p.arrows == 2 or p.connect != "single" else None)
            if p.connect == "single":
                if p.to is None and p.arrows == 2:
                    pairs = [(f, f) for f in frm]
                else:
                    k = max(len(frm), len(to))
                    pairs = [(frm[i % len(frm)], to[i % len(to)]) for i in range(k)]
            elif p.connect == "all.pairs":
                pairs = [(f, t) for f in frm for t in to]
            elif p.connect == "unique.bivariate":
                tl = to if p.to is not None else frm
                pairs = [(frm[i], tl[j]) for i in range(len(frm)) for j in range(i + 1, len(tl))]
            else:
                raise ValueError(p.connect)
            bc = lambda x, k, n: (x[k % len(x)] if isinstance(x, (list, tuple, np.ndarray)) else x)
            for k, (f, t) in enumerate(pairs):
                free = bool(bc(p.free, k, len(pairs)))
                val = bc(p.values, k, len(pairs))
                lab = bc(p.labels, k, len(pairs))
                lb = bc(p.lbound, k, len(pairs))
                if f == "one":
                    mats, cells = [self.M], [(0, idx[t])]
                elif p.arrows == 1:
                    mats, cells = [self.A], [(idx[t], idx[f])]
                else:
                    mats, cells = [self.S, self.S], [(idx[t], idx[f]), (idx[f], idx[t])]
                for mat, (r, c) in zip(mats, cells):
                    mat["free"][r, c] = free
                    if val is not None and not (isinstance(val, float) and np.isnan(val)):
                        mat["values"][r, c] = float(val)
                    if lab is not None:
                        mat["labels"][r, c] = lab
                    if lb is not None:
                        mat["lbound"][r, c] = float(lb)
        return self

    # -- free parameter table in OpenMx order (A, S, M column-major, then thresholds) ---
    def parameters(self):
        """Returns (names, starts, lbounds, locations) where locations[k] is a list of
        (matrix, row, col) cells sharing parameter k."""
        names, starts, lbs, locs = [], [], [], []
        by_name = {}

        def add(mat_name, r, c, lab, val, lb):
            nm = lab if lab is not None else f"{self.name}.{mat_name}[{r + 1},{c + 1}]"
            if nm not in by_name:
                by_name[nm] = len(names)
                names.append(nm)
                starts.append(val)
                lbs.append(lb)
                locs.append([])
            locs[by_name[nm]].append((mat_name, r, c))

        n = len(self.vars)
        for c in range(n):
            for r in range(n):
                if self.A["free"][r, c]:
                    add("A", r, c, self.A["labels"][r, c], self.A["values"][r, c], self.A["lbound"][r, c])
        for c in range(n):
            for r in range(c, n):  # symmetric: lower triangle names the parameter, both cells carry it
                if self.S["free"][r, c]:
                    add("S", r, c, self.S["labels"][r, c], self.S["values"][r, c], self.S["lbound"][r, c])
                    if r != c:
                        locs[by_name[names[-1] if self.S["labels"][r, c] is None else self.S["labels"][r, c]]].append(("S", c, r))
        if self.hasMeans:
            for c in range(n):
                if self.M["free"][0, c]:
                    add("M", 0, c, self.M["labels"][0, c], self.M["values"][0, c], self.M["lbound"][0, c])
        for item, th in self.thresholds.items():
            for k in range(len(th["values"])):
                if th["free"][k]:
                    add("T:" + item, k, 0, th["labels"][k], th["values"][k], np.nan)
        return names, np.array(starts, float), np.array(lbs, float), locs

    def ordinal_items(self):
        return [m for m in self.manifestVars if m in self.thresholds]

    def definition_means(self):
        """{latent covariate -> data column} for M cells labelled 'data.<col>' (R/model.R:407)."""
        out = {}
        for c, v in enumerate(self.vars):
            lab = self.M["labels"][0, c]
            if isinstance(lab, str) and lab.startswith("data."):
                out[v] = lab[5:]
        return out


# ------------------------------------------------------------------------------- helpers

def setupThresholds(model):  # R/model.R:346-365
    pheno = model.data
    for m1 in [m for m in model.manifestVars if m != "snp"]:
        if m1 in pheno and _is_factor(pheno[m1]):
            nth = len(pheno[m1].cat.categories) - 1
            if nth < 1:
                continue
            model.thresholds[m1] = dict(
                values=list(mxNormalQuantiles(nth, sd=np.sqrt(2.0))),
                labels=[f"{m1}_Thr_{k}" for k in range(1, nth + 1)],
                free=[True] * nth)
    return model


def setupExogenousCovariates(model, covariates, itemNames):  # R/model.R:403-418
    if not covariates:
        return model
    model.add_paths([
        mxPath("one", covariates, free=False, labels=["data." + c for c in covariates]),
        mxPath(covariates, itemNames, connect="all.pairs",
               labels=[f"{c}_to_{i}" for c in covariates for i in itemNames])])
    exo = pd.DataFrame(False, index=model.manifestVars, columns=list(covariates))
    exo.loc[list(itemNames), list(covariates)] = True
    model.exoFree = exo
    model.exoPred = list(covariates)
    return model


def setupData(phenoData, gxe, customMinMAF, minMAF, fitfun):  # R/model.R:421-437
    import warnings
    if customMinMAF and fitfun != "WLS":
        warnings.warn("minMAF is ignored when fitfun != 'WLS'")
    phenoData = phenoData.copy()
    for v1 in gxe or []:
        phenoData["snp_" + v1] = 0.0  # placeholder, filled per SNP with snp * E
    return phenoData, calcMinVar(minMAF)


def addPlaceholderSNP(phenoData):  # R/model.R:440-448
    import warnings
    phenoData = pd.DataFrame(phenoData).copy()
    if "snp" in phenoData:
        warnings.warn("Data already contains placeholder data for the 'snp' column. "
                      "This is okay for testing, but generally not recommended")
    else:
        phenoData["snp"] = np.random.default_rng(0).binomial(2, .5, len(phenoData)).astype(float)
    return phenoData


def endogenousSNPpath(pred, depVar):  # R/model.R:450-457
    return [mxPath("one", pred, labels=[p + "Mean" for p in pred]),
            # connect="single": from/to paired elementwise with R recycling, labels likewise
            mxPath(pred, depVar, values=0,
                   labels=[f"{pred[k % len(pred)]}_to_{depVar[k % len(depVar)]}"
                           for k in range(max(len(pred), len(depVar)))]),
            mxPath(pred, arrows=2, values=1, lbound=0.001, labels=[p + "_res" for p in pred])]


def endogenousCovariatePaths(phenoData, covariates, depVar):  # R/model.R:459-479
    paths = []
    for c1 in covariates or []:
        if _is_factor(phenoData[c1]):
            paths.append(mxPath(c1, arrows=2, values=1, free=False))
        else:
            paths += [mxPath("one", c1, labels=c1 + "_mean"),
                      mxPath(c1, arrows=2, values=1, lbound=0.001, labels=c1 + "_var")]
        paths.append(mxPath(c1, depVar, connect="all.pairs", labels=[f"{c1}_to_{d}" for d in depVar]))
    return paths


def postprocessModel(model, indicators, exogenous):  # R/model.R:482-512
    import warnings
    model = setupThresholds(model)
    # R: typeof(x) == 'integer' -- true for integer vectors and for factors alike
    isInt = [c for c in model.data.columns
             if pd.api.types.is_integer_dtype(model.data[c].dtype) or _is_factor(model.data[c])]
    if not exogenous and not model.thresholds and model.fitfun == "WLS":
        if isInt:
            warnings.warn("Cumulants method prevented by integer observed columns: " +
                          ", ".join(f"'{c}'" for c in isInt) + "; Remove these columns from your observed "
                          "data for improved estimation accuracy")
        else:
            model.continuousType = "cumulants"
            model.hasMeans = False  # discard model means
    else:
        model = setupExogenousCovariates(model, exogenous, indicators)
    return model


def _aslist(x):
    if x is None:
        return []
    return [x] if isinstance(x, str) else list(x)


def _check_dots(dots):
    if dots:
        raise ValueError("Rejected are any values passed in the '...' argument")


def _fitfun(fitfun):
    if isinstance(fitfun, (list, tuple)):
        fitfun = fitfun[0]
    if fitfun not in ("WLS", "ML"):
        raise ValueError(f"'arg' should be one of 'WLS', 'ML'")
    return fitfun


_MINMAF_DEFAULT = object()


def _model(name, manifest, latents, paths, phenoData, gxe, minMAF, fitfun):
    custom = minMAF is not _MINMAF_DEFAULT
    data, minVar = setupData(phenoData, gxe, custom, 0.01 if not custom else minMAF, fitfun)
    m = RAMModel(name=name, manifestVars=list(manifest), latentVars=list(latents), data=data,
                 fitfun=fitfun, minVariance=minVar, gxe=list(gxe or []))
    return m.add_paths(paths)


# ------------------------------------------------------------------------------ builders

def buildItem(phenoData, depVar, covariates=None, *dots, fitfun="WLS", minMAF=_MINMAF_DEFAULT,
              gxe=None, exogenous=None, pred="snp"):
    """R/model.R:548-598."""
    import warnings
    _check_dots(dots)
    fitfun = _fitfun(fitfun)
    depVar, covariates, gxe = _aslist(depVar), _aslist(covariates), _aslist(gxe)
    phenoData = addPlaceholderSNP(phenoData)
    keep = ["snp"] + depVar + covariates + gxe
    phenoData = phenoData[[c for c in phenoData.columns if c in keep]]
    fac = [_is_factor(phenoData[d]) for d in depVar]
    if exogenous is None:
        exogenous = not (fitfun == "WLS" and not any(fac))
    pred = _aslist(pred) + ["snp_" + g for g in gxe]
    manifest, latents, endo = list(depVar), [], []
    if not exogenous:
        manifest += pred + covariates
        endo = covariates
    else:
        if minMAF is not _MINMAF_DEFAULT:
            warnings.warn("minMAF is ignored when SNP is not an observed variable")
        latents = pred + covariates
    paths = endogenousCovariatePaths(phenoData, endo, depVar)
    if not exogenous:
        paths += endogenousSNPpath(pred, depVar)
    notfac = [not f for f in fac]
    paths += [mxPath(depVar, arrows=2, values=1, lbound=0.001, free=notfac, labels=[d + "_res" for d in depVar]),
              mxPath(depVar, arrows=2, values=0, connect="unique.bivariate"),
              mxPath("one", depVar, free=notfac, values=0, labels=[d + "Mean" for d in depVar])]
    model = _model("OneItem", manifest, latents, paths, phenoData, gxe, minMAF, fitfun)
    return postprocessModel(model, depVar, latents)


def _factor_builder(name, phenoData, itemNames, factors, snp_targets, covariates, fitfun, minMAF, gxe,
                    exogenous, pred, loading_paths, cov_target):
    fitfun = _fitfun(fitfun)
    covariates, gxe = _aslist(covariates), _aslist(gxe)
    phenoData = addPlaceholderSNP(phenoData)
    fac = [_is_factor(phenoData[i]) for i in itemNames]
    if exogenous is None:
        exogenous = defaultExogenous
    pred = _aslist(pred)
    manifest = pred + list(itemNames) + ["snp_" + g for g in gxe]
    latents, exoPred, endo = list(factors), [], []
    if exogenous:
        latents += covariates
        exoPred = covariates
    else:
        manifest += covariates
        endo = covariates
    pred = pred + ["snp_" + g for g in gxe]
    notfac = [not f for f in fac]
    paths = endogenousSNPpath(pred, snp_targets) + endogenousCovariatePaths(phenoData, endo, cov_target)
    paths += loading_paths
    paths += [mxPath(itemNames, arrows=2, values=1, lbound=0.001, free=notfac,
                     labels=[i + "_res" for i in itemNames]),
              mxPath(factors, arrows=2, free=False, values=1.0, labels="facRes"),
              mxPath("one", itemNames, free=notfac, values=0, labels=[i + "Mean" for i in itemNames])]
    model = _model(name, manifest, latents, paths, phenoData, gxe, minMAF, fitfun)
    return postprocessModel(model, list(itemNames), exoPred)


def buildOneFac(phenoData, itemNames, covariates=None, *dots, fitfun="WLS", minMAF=_MINMAF_DEFAULT,
                gxe=None, exogenous=None, pred="snp"):
    """R/model.R:635-679."""
    _check_dots(dots)
    itemNames = _aslist(itemNames)
    load = [mxPath("F", itemNames, values=1, free=True, labels=["lambda_" + i for i in itemNames])]
    return _factor_builder("OneFac", phenoData, itemNames, ["F"], ["F"], covariates, fitfun, minMAF, gxe,
                           exogenous, pred, load, ["F"])


def buildOneFacRes(phenoData, itemNames, factor=False, res=None, covariates=None, *dots, fitfun="WLS",
                   minMAF=_MINMAF_DEFAULT, gxe=None, exogenous=None, pred="snp"):
    """R/model.R:711-758."""
    _check_dots(dots)
    itemNames = _aslist(itemNames)
    res = itemNames if res is None else _aslist(res)
    depVar = (["F"] if factor else []) + res
    load = [mxPath("F", itemNames, values=1, free=True, labels=["lambda_" + i for i in itemNames])]
    return _factor_builder("OneFacRes", phenoData, itemNames, ["F"], depVar, covariates, fitfun, minMAF, gxe,
                           exogenous, pred, load, ["F"])


def buildTwoFac(phenoData, F1itemNames, F2itemNames, covariates=None, *dots, fitfun="WLS",
                minMAF=_MINMAF_DEFAULT, gxe=None, exogenous=None, pred="snp"):
    """R/model.R:786-834."""
    _check_dots(dots)
    F1, F2 = _aslist(F1itemNames), _aslist(F2itemNames)
    itemNames = F1 + [i for i in F2 if i not in F1]
    load = [mxPath("F1", F1, values=1, labels=["F1_lambda_" + i for i in F1]),
            mxPath("F2", F2, values=1, labels=["F2_lambda_" + i for i in F2]),
            mxPath("F1", "F2", arrows=2, free=True, values=.3, labels="facCov")]
    return _factor_builder("TwoFac", phenoData, itemNames, ["F1", "F2"], ["F1", "F2"], covariates, fitfun,
                           minMAF, gxe, exogenous, pred, load, ["F1", "F2"])


def buildOneItem(*args, **kw):  # R/model.R:602-608 (deprecated alias)
    import warnings
    warnings.warn("buildOneItem() is deprecated, use buildItem()", DeprecationWarning)
    return buildItem(*args, **kw)


# ------------------------------------------------------------------- flattened model spec

def flatten(model):
    """Flatten a RAMModel into plain arrays: the model half of the C-ABI spec
    (include/gwsem_b200.h, gwsem_model_spec) and the input of oracle/fit_oracle.py.

    Cell codes: matrix 0 = A, 1 = S, 2 = M.  `par` = free-parameter index or -1.
    """
    names, starts, lbs, locs = model.parameters()
    nv, nm = len(model.vars), len(model.manifestVars)
    cells = []  # (matrix, row, col, par, value)
    mats = [("A", model.A, 0), ("S", model.S, 1)] + ([("M", model.M, 2)] if model.hasMeans else [])
    cell_par = {}
    for k, ll in enumerate(locs):
        for (mn, r, c) in ll:
            cell_par[(mn, r, c)] = k
    for mn, mat, code in mats:
        rows = range(1) if mn == "M" else range(nv)
        for r in rows:
            for c in range(nv):
                par = cell_par.get((mn, r, c), -1)
                val = mat["values"][r, c]
                lab = mat["labels"][r, c]
                if par < 0 and isinstance(lab, str) and lab.startswith("data."):
                    continue  # definition variable: enters through the slopes, mean handled as 0
                if par >= 0 or val != 0.0:
                    cells.append((code, r, c, par, float(val)))
    thr = []  # (manifest index, k, par, value)
    for item, th in model.thresholds.items():
        mi = model.manifestVars.index(item)
        for k in range(len(th["values"])):
            thr.append((mi, k, cell_par.get(("T:" + item, k, 0), -1), float(th["values"][k])))
    return dict(name=model.name, nv=nv, nm=nm, manifest=list(model.manifestVars), latent=list(model.latentVars),
                par_names=names, par_start=starts, par_lbound=lbs,
                cells=np.array(cells, float).reshape(-1, 5), thresholds=np.array(thr, float).reshape(-1, 4),
                has_means=bool(model.hasMeans), fitfun=model.fitfun, continuous_type=model.continuousType,
                min_variance=float(model.minVariance), gxe=list(model.gxe),
                exo_pred=list(model.exoPred),
                exo_free=None if model.exoFree is None else model.exoFree.to_numpy(bool))


# ---------------------------------------------------------------------------- multigroup helpers
# tests/testthat/test-multigroup.R:15-36 assembles a multigroup model from OpenMx pieces; these are
# the minimal counterparts: mxRename, omxSetParameters (relabelling only), mxModel(name, g1, g2,
# mxFitFunctionMultigroup(...)).

def mxRename(model, newname):
    """A copy of `model` under another name."""
    import copy
    m = copy.deepcopy(model)
    m.name = newname
    return m


def coefNames(model):
    """names(coef(model)): the free-parameter labels in the order the log uses."""
    return list(flatten(model)["par_names"])


def omxSetParameters(model, labels, newlabels):
    """Relabel free parameters (the only use the reference's tests make of omxSetParameters)."""
    import copy
    labels, newlabels = list(labels), list(newlabels)
    if len(labels) != len(newlabels):
        raise ValueError("labels and newlabels differ in length")
    ren = dict(zip(labels, newlabels))
    m = copy.deepcopy(model)
    for mat in (m.A, m.S, m.M):
        lab = mat["labels"]
        for idx in np.ndindex(lab.shape):
            if lab[idx] in ren:
                lab[idx] = ren[lab[idx]]
    for thr in m.thresholds.values():
        thr["labels"] = [ren.get(l, l) for l in thr["labels"]]
    return m


class MultigroupModel:
    """mxModel(name, submodels..., mxFitFunctionMultigroup(names)): the fit is the sum of the groups' fits."""

    def __init__(self, name, *submodels):
        self.name = name
        self.submodels = list(submodels)
        names = [m.name for m in self.submodels]
        if len(set(names)) != len(names):
            raise ValueError("sub-models must have distinct names")
