"""Result readers: R/report.R mirrored on pandas (loadResults, signif, signifGxE, isSuspicious)."""
import re

import numpy as np
import pandas as pd
from scipy.stats import norm


def isSuspicious(result, pars=None):
    """R/report.R:32-48."""
    if pars is None:
        pars = result.attrs.get("focus", [])
    if isinstance(pars, str):
        pars = [pars]
    catch = result["catch1"]
    mask = (catch.notna() & (catch != "")) | (result["statusCode"] != "OK")
    for p1 in pars:
        if p1 not in result:
            import warnings
            warnings.warn(f"{p1} not found; ignoring")
            continue
        mask = mask | result[p1].isna()
        pvar = f"V{p1}:{p1}"
        if pvar in result:
            mask = mask | result[pvar].isna()
    return mask


def loadResults(path, focus, *dots, extraColumns=()):
    """R/report.R:77-116: read the checkpoint log(s), selecting columns by name."""
    if dots:
        raise ValueError("Rejected are any values passed in the '...' argument")
    paths = [path] if isinstance(path, str) else list(path)
    focus = [focus] if isinstance(focus, str) else list(focus)
    header = pd.read_csv(paths[0], sep="\t", nrows=0, quoting=3).columns
    sel = ["MxComputeLoop1", "CHR", "BP", "SNP", "A1", "A2", "statusCode", "catch1"]
    sel += focus + [f"V{f}:{f}" for f in focus] + list(extraColumns)
    for f in focus:
        m = re.match(r"^snp_(\w+)_to_(\w+)$", f)
        if not m:
            continue
        main = f"snp_to_{m.group(2)}"
        cov = [c for c in (f"V{f}:{main}", f"V{main}:{f}") if c in header]
        if len(cov) != 1:
            raise ValueError(f"Can't find parameter covariance between {main} and {f}")
        for c in (main, f"V{main}:{main}", cov[0]):
            if c not in sel:
                sel.append(c)
    got = []
    for p1 in paths:
        d = pd.read_csv(p1, sep="\t", quoting=3, keep_default_na=True, na_values=["NA"],
                        dtype={"catch1": str, "statusCode": str, "SNP": str, "A1": str, "A2": str, "CHR": str})
        d["catch1"] = d["catch1"].fillna("")
        got.append(d[[c for c in sel if c in d.columns]])
    return pd.concat(got, ignore_index=True)


def loadSuspicious(path, focus, *dots, extraColumns=(), signAdj=None, moderatorLevel=None):
    """R/report.R:139-143: removed in the reference since 2.0.6 (deprecate_stop); same behaviour here."""
    raise RuntimeError("`loadSuspicious()` was deprecated in gwsem 2.0.6 and is now defunct. "
                       "Use r1 = loadResults(path, focus); r1[isSuspicious(r1, focus)] instead.")


def signif(result, focus, signAdj=None):
    """R/report.R:174-190: Z = est / sqrt(V), two-sided normal P."""
    if not isinstance(focus, str):
        raise ValueError("Can only do one at a time")
    result = result.copy()
    if signAdj is not None:
        if signAdj not in result:
            raise ValueError(f"signAdj= {signAdj} but no such column found in result")
        result[focus] = result[focus] * np.sign(result[signAdj])
    with np.errstate(invalid="ignore", divide="ignore"):
        result["Z"] = result[focus] / np.sqrt(result[f"V{focus}:{focus}"])
    result["P"] = 2 * norm.cdf(-np.abs(result["Z"]))
    result.attrs["focus"] = focus
    return result


def signifGxE(result, focus, level):
    """R/report.R:203-228: marginal effect of the SNP at moderator level `level`."""
    m = re.match(r"^snp_(\w+)_to_(\w+)$", focus)
    if not m:
        raise ValueError("focus must be of the form snp_XXX_to_F")
    main = f"snp_to_{m.group(2)}"
    cov = [c for c in (f"V{focus}:{main}", f"V{main}:{focus}") if c in result]
    result = result.copy()
    tmp = result[main] + level * result[focus]
    with np.errstate(invalid="ignore"):
        se = np.sqrt(result[f"V{main}:{main}"] + level ** 2 * result[f"V{focus}:{focus}"] + 2 * level * result[cov[0]])
    result["ME"], result["ME.SE"] = tmp, se
    result["Z"] = tmp / se
    result["P"] = 2 * norm.cdf(-np.abs(result["Z"]))
    result.attrs["focus"] = focus
    return result
