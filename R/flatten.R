# R-side glue between the unchanged gwsem model builders (R/model.R:317-834 of the reference) and
# libgwsem_b200: everything GWAS() needs to hand the per-SNP loop to the GPU engine in one .Call
# (src/gwsem_call.c).  OpenMx is used only as the *description* of the model -- its R constructors
# build the MxModel; mxRun() is never called.
#
#   flattenModel(model)      the RAM matrices, thresholds, exogenous covariates and fit function as flat vectors
#   manifestColumns(model)   the observed columns behind the manifest variables
#   jobList(...)             SNP selection, log path, labels, context file, vcov filter
#   GWAS(...)                drop-in for R/model.R:305-315
#   numAvailableRecords(...) drop-in for R/model.R:210-224
#
# gwsem_b200/model.py (flatten) and gwsem_b200/gwas.py (call_arguments) are the same code in Python; the C
# harness `make -C src check` feeds their output through the .Call layer, because R is not installed where this
# package is developed.  Field names and orders below and there must stay identical.

# free parameters in OpenMx's order: A then S (lower triangle, column-major) then M, then thresholds
.gwsemParameterTable <- function(model) {
  vars <- c(model$manifestVars, model$latentVars)
  names <- character(0); starts <- numeric(0); lbs <- numeric(0)
  cellPar <- list()
  add <- function(mat, r, c, lab, val, lb) {
    nm <- if (is.na(lab)) paste0(model$name, ".", mat, "[", r, ",", c, "]") else lab
    k <- match(nm, names)
    if (is.na(k)) {
      names <<- c(names, nm); starts <<- c(starts, val); lbs <<- c(lbs, lb)
      k <- length(names)
    }
    cellPar[[paste(mat, r, c)]] <<- k
    k
  }
  A <- model$A; S <- model$S; M <- model$M
  n <- length(vars)
  for (c in seq_len(n)) for (r in seq_len(n))
    if (A$free[r, c]) add("A", r, c, A$labels[r, c], A$values[r, c], A$lbound[r, c])
  for (c in seq_len(n)) for (r in c:n)
    if (S$free[r, c]) {
      k <- add("S", r, c, S$labels[r, c], S$values[r, c], S$lbound[r, c])
      if (r != c) cellPar[[paste("S", c, r)]] <- k
    }
  if (!is.null(M)) for (c in seq_len(n))
    if (M$free[1, c]) add("M", 1, c, M$labels[1, c], M$values[1, c], M$lbound[1, c])
  thr <- model$Thresholds
  if (!is.null(thr)) for (c in seq_len(ncol(thr$values))) for (r in seq_len(nrow(thr$values)))
    if (thr$free[r, c]) add(paste0("T:", colnames(thr$values)[c]), r, 1, thr$labels[r, c], thr$values[r, c], NA_real_)
  list(names = names, starts = starts, lbound = lbs, cellPar = cellPar)
}

#' Flatten an MxModel built by buildItem / buildOneFac / buildOneFacRes / buildTwoFac
#'
#' @param model an MxModel of type RAM with raw data and a WLS or ML fit function
#' @return a named list, the `spec` argument of .Call("gwsem_run", ...)
flattenModel <- function(model) {
  if (!is(model$expectation, "MxExpectationRAM")) stop("Only RAM models are supported")
  manifest <- model$manifestVars
  vars <- c(manifest, model$latentVars)
  nv <- length(vars); nm <- length(manifest)
  pt <- .gwsemParameterTable(model)
  observed <- model$data$observed
  gxe <- character(0)
  if (length(model$data$algebra))
    gxe <- sub("^snp_", "", sapply(model$data$algebra, function(x) model[[x]]$.dimnames[[2]]))

  # cells: matrix code (0 A, 1 S, 2 M), 0-based row, 0-based column, 0-based parameter or -1, value
  cells <- list()
  one <- function(code, mat, mname, rows) {
    for (r in rows) for (c in seq_len(nv)) {
      par <- pt$cellPar[[paste(mname, r, c)]]
      val <- mat$values[r, c]
      lab <- mat$labels[r, c]
      # definition variables ('data.<column>' on a fixed mean) enter through the slopes; their mean is handled as 0
      if (is.null(par) && !is.na(lab) && startsWith(lab, "data.")) next
      if (!is.null(par) || val != 0)
        cells[[length(cells) + 1]] <<- c(code, r - 1, c - 1, if (is.null(par)) -1 else par - 1, val)
    }
  }
  one(0, model$A, "A", seq_len(nv))
  one(1, model$S, "S", seq_len(nv))
  hasMeans <- !is.null(model$M)
  if (hasMeans) one(2, model$M, "M", 1)
  cells <- if (length(cells)) do.call(rbind, cells) else matrix(0, 0, 5)

  # thresholds: 0-based manifest index, 0-based k, 0-based parameter or -1, value
  thr <- matrix(0, 0, 4)
  nCat <- integer(nm)
  tm <- model$Thresholds
  if (!is.null(tm)) for (item in colnames(tm$values)) {
    mi <- match(item, manifest)
    nth <- nlevels(observed[[item]]) - 1L
    nCat[mi] <- nth + 1L
    for (k in seq_len(nth)) {
      par <- pt$cellPar[[paste(paste0("T:", item), k, 1)]]
      thr <- rbind(thr, c(mi - 1, k - 1, if (is.null(par)) -1 else par - 1, tm$values[k, item]))
    }
  }

  # manifest kinds: 0 data column, 1 snp, 2 snp * moderator (the gxe algebra columns, R/model.R:427-434)
  kind <- integer(nm)
  kind[manifest == "snp"] <- 1L
  kind[manifest %in% paste0("snp_", gxe)] <- 2L

  # exogenous covariates: latent variables whose mean is a definition variable (R/model.R:403-418)
  exoPred <- character(0)
  if (hasMeans) {
    lab <- model$M$labels[1, ]
    exoPred <- names(lab)[!is.na(lab) & startsWith(lab, "data.")]
  }
  exoColumn <- sub("^data\\.", "", model$M$labels[1, exoPred])
  exoKind <- integer(length(exoPred))
  exo <- vector("list", length(exoPred))
  for (k in seq_along(exoPred)) {
    col <- exoColumn[k]
    if (col == "snp") { exoKind[k] <- 1L; next }               # the genotype itself: NULL column
    if (col %in% paste0("snp_", gxe)) { exoKind[k] <- 2L; col <- sub("^snp_", "", col) }
    exo[[k]] <- as.numeric(observed[[col]])
  }
  exoFree <- model$data$exoFree
  xf <- matrix(FALSE, nm, length(exoPred), dimnames = list(manifest, exoPred))
  if (!is.null(exoFree)) xf[rownames(exoFree), colnames(exoFree)] <- exoFree[rownames(exoFree), colnames(exoFree)]

  wls <- is(model$fitfunction, "MxFitFunctionWLS")
  fit <- if (!wls) 2L else if (identical(model$fitfunction$continuousType, "cumulants")) 0L else 1L
  list(n_vars = nv, n_manifest = nm, n_params = length(pt$names), n_cells = nrow(cells),
       cells = as.numeric(t(cells)),                 # row major
       par_start = as.numeric(pt$starts), par_lbound = as.numeric(pt$lbound),
       fit = fit, min_variance = as.numeric(model$data$minVariance), n_rows = nrow(observed),
       manifest_kind = kind, n_categories = nCat,
       exo = exo, exo_latent = as.integer(match(exoPred, vars) - 1L), exo_kind = exoKind,
       exo_free = as.raw(t(xf)),                      # n_manifest x n_exo, row major
       n_thresholds = nrow(thr), thresholds = as.numeric(t(thr)),
       par_names = pt$names, gxe = gxe)
}

#' The observed columns behind the manifest variables: doubles; ordinal items as 0-based codes with NA
#' kept; NULL for snp; the moderator for a gxe product column
manifestColumns <- function(model, spec = flattenModel(model)) {
  observed <- model$data$observed
  lapply(model$manifestVars, function(v) {
    if (v == "snp") return(NULL)
    col <- if (v %in% paste0("snp_", spec$gxe)) sub("^snp_", "", v) else v
    x <- observed[[col]]
    if (is.factor(x)) as.numeric(as.integer(x) - 1L) else as.numeric(x)
  })
}

#' What prepareComputePlan (R/model.R:122-195) decides, as data
jobList <- function(model, out, SNP, startFrom, psd, spec = flattenModel(model), device = 0L, batchSnps = 0L) {
  ext <- psd$snpFileExt[1]; stem <- psd$stem[1]
  context <- switch(ext, pgen = paste(stem, "pvar", sep = "."), bed = paste(stem, "bim", sep = "."), NULL)
  # off-diagonal vcov entries needed by signifGxE (R/model.R:176-185)
  pairs <- NULL
  if (length(spec$gxe)) {
    gxeCols <- paste0("snp_", spec$gxe)
    outcome <- model$A$free[, "snp"]
    outcome <- names(outcome)[outcome]
    want <- c(paste0("snp_to_", outcome),
              apply(expand.grid(c(gxeCols, spec$gxe), "_to_", outcome), 1, paste0, collapse = ""))
    idx <- match(intersect(want, spec$par_names), spec$par_names) - 1L
    if (length(idx) > 1) pairs <- as.integer(combn(idx, 2))
  }
  list(device = as.integer(device), snp = if (is.null(SNP)) NULL else as.integer(SNP),
       start_from = as.integer(startFrom), out = path.expand(out), par_names = spec$par_names,
       context_path = context, vcov_pairs = pairs, batch_snps = as.integer(batchSnps),
       model_name = model$name, manifest_names = model$manifestVars)
}

#' Run a genome-wide association study with the specified model (drop-in for R/model.R:305-315)
GWAS <- function(model, snpData, out = "out.log", ..., SNP = NULL, startFrom = 1L, rowFilter = NULL) {
  if (length(list(...)) > 0) stop("Rejected are any values passed in the '...' argument")
  psd <- processSnpData(snpData)                          # unchanged, R/model.R:24-45
  if (length(snpData) > 1) stop("one genotype file per call (independent groups: one call per group)")
  noMatch <- is.na(match(names(rowFilter), model$name))
  if (any(noMatch)) stop(paste("Cannot find model associated with rowFilter:", omxQuotes(names(rowFilter)[noMatch])))
  rf <- rowFilter[[model$name]]
  if (!is.null(rf) && sum(!rf) != nrow(model$data$observed))
    stop(paste("rowFilter keeps", sum(!rf), "rows, which does not match the number of rows of observed data"))
  spec <- flattenModel(model)
  rows <- .Call("gwsem_run", spec[setdiff(names(spec), c("par_names", "gxe"))], manifestColumns(model, spec),
                path.expand(snpData), psd$method[1], if (is.null(rf)) NULL else as.logical(rf),
                jobList(model, out, SNP, startFrom, psd, spec), PACKAGE = "gwsem")
  message(paste("Done. See", omxQuotes(out), "for results"))
  invisible(model)
}

#' Number of records available in each genotype file (drop-in for R/model.R:210-224)
numAvailableRecords <- function(snpData) {
  psd <- processSnpData(snpData)
  result <- mapply(function(path, method) .Call("gwsem_num_records", path.expand(path), method, -1L, PACKAGE = "gwsem"),
                   snpData, psd$method)
  names(result) <- snpData
  result
}
