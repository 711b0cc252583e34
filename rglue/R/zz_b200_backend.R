# zz_b200_backend.R -- re-points the INTERNAL hot-path functions of funGp at libfungp_b200 through .Call.
# Collated last (DESCRIPTION Collate), so the definitions that keep a reference name (dimReduction, getOut_loocv) replace
# the R bodies of the same names; the *_B200 helpers are called from the re-pointed bodies of fgpm(), predict.fgpm,
# simulate.fgpm, update.fgpm and run_ACO -- see rglue/patches/*.diff, which wire them in.  The exported API (fgpm,
# predict, simulate, update, fgpm_factory) and every signature visible to users is unchanged; no S4 slot is added (the
# device handle travels as an attribute of the preMats list, so objects saved by the unpatched package still load).
# NAMESPACE gains:  useDynLib(funGp, .registration = TRUE)
# This file has not been run: R is not installed in the build image (see INTEGRATION.md).

.b200 <- new.env(parent = emptyenv())
.KER  <- c(gauss = 1L, matern5_2 = 2L, matern3_2 = 3L)      # level order of R/1_Xfgpm_Class.R:457-459
.DIST <- c(L2_bygroup = 1L, L2_byindex = 2L)
.BAS  <- c("B-splines" = 1L, PCA = 2L)

b200_ctx <- function() {
  if (is.null(.b200$ctx)) {
    n <- getOption("funGp.gpus", as.integer(Sys.getenv("FUNGP_GPUS", "0")))
    .b200$ctx <- .Call(fgpR_ctx_create, as.integer(n))
  }
  .b200$ctx
}

# dimReduction (R/7_dimRedFunctions.R:7-29): same return value list(basis, coefs, J)
dimReduction <- function(fIn, df, fpDims, methvec) {
  basis <- coefs <- J <- list()
  for (i in seq_len(df)) {
    bt <- if (fpDims[i] > 0) .BAS[[methvec[i]]] else 1L
    r <- .Call(fgpR_project, b200_ctx(), fIn[[i]], as.integer(fpDims[i]), bt, NULL)
    basis[[i]] <- r[[1]]; J[[i]] <- r[[2]]; coefs[[i]] <- r[[3]]
  }
  list(basis = basis, coefs = coefs, J = J)
}

# setDistMatrix_S / setDistMatrix_F (R/7_distanceFunctions.R:3-22) are replaced by ONE device problem handle that
# travels where sMs / fMs travelled (fgpm(): R/1_fgpm_Class.R:404-407).  Code that really needs a distance matrix
# gets it with b200_dist(prob, l).
b200_problem <- function(sIn, coefs, J, f_disType, sOut)
  .Call(fgpR_problem_create, b200_ctx(), sIn, coefs, J, unname(.DIST[f_disType]), as.numeric(sOut))
b200_dist <- function(prob, l) .Call(fgpR_dist_matrix, prob, as.integer(l))

# The handle of a fitted model.  It is an attribute of model@preMats (a plain list slot): copies of the model share it,
# save()/load() turns it into a nil pointer.  A model without a live handle (loaded from disk, or built by the unpatched
# package) gets one rebuilt from its slots: the problem from sIn / f_proj@coefs / crossprod(basis) / sOut, and the stored
# preMats handed back with fgp_problem_restore -- no assembly, no factorisation (SURVEY.md 8 f4).  Rebuilt handles are
# kept in a small cache keyed by the model's numbers, because R's value semantics give the caller's object no way to
# receive them.
b200_handle <- function(model) {
  p <- attr(model@preMats, "b200prob")
  if (!is.null(p) && .Call(fgpR_ptr_valid, p)) return(p)
  z <- model@preMats$LInvY
  key <- sprintf("%d:%.17g:%.17g:%.17g", length(z), z[1], sum(z), model@kern@varHyp)
  p <- .b200$cache[[key]]
  if (!is.null(p) && .Call(fgpR_ptr_valid, p)) return(p)
  coefs <- if (model@df > 0) model@f_proj@coefs else list()
  J <- if (model@df > 0) lapply(model@f_proj@basis, crossprod) else list()
  p <- b200_problem(if (model@ds > 0) model@sIn else NULL, coefs, J, model@kern@f_disType, model@sOut)
  .Call(fgpR_restore, p, .KER[[model@kern@kerType]], model@nugget, model@kern@varHyp,
        as.numeric(c(model@kern@s_lsHyps, model@kern@f_lsHyps)), model@preMats$L, as.numeric(z))
  if (is.null(.b200$cache)) .b200$cache <- list()
  if (length(.b200$cache) >= 8) .b200$cache <- .b200$cache[-1]       # a handful of models at most
  .b200$cache[[key]] <- p
  p
}

# setBounds_* (R/3_training_SF.R:57-74): lower 1e-10, upper 2 * max(D_l)
setBounds_B200 <- function(prob, free) {
  mx <- .Call(fgpR_dist_max, prob)[free]
  rbind(rep(10^-10, length(mx)), 2 * mx)
}

# negLogLik_funGp_SF (R/3_training_SF.R:231-258) -- also accepts a d x B matrix of points (one launch sequence)
negLogLik_funGp_B200 <- function(thetas, prob, kerType, var.known, full_theta, nugget) {
  r <- .Call(fgpR_nll_batch, prob, .KER[[kerType]], nugget, full_theta(thetas), var.known)
  if (any(r[[3]] > 0L))                                  # same condition object base::chol raises at :249
    stop(sprintf("the leading minor of order %d is not positive definite", r[[3]][which(r[[3]] > 0L)[1]]))
  r[[1]]
}

# setSPoints_* (R/3_training_SF.R:83-108): the runif draw is unchanged (same RNG stream), the n.presample
# likelihoods come back from one batched call instead of apply(...)
setSPoints_B200 <- function(bnds, prob, kerType, var.known, full_theta, n.starts, n.presample, spoints.usr, nugget) {
  ll <- bnds[1, ]; ul <- bnds[2, ]; n.ls <- ncol(bnds)
  allspoints <- ll + matrix(runif(n.ls * n.presample), nrow = n.ls) * (ul - ll)
  fitvec <- negLogLik_funGp_B200(allspoints, prob, kerType, var.known, full_theta, nugget)
  spoints <- matrix(allspoints[, order(fitvec)[1:n.starts]], ncol = n.starts)
  if (!is.null(spoints.usr)) {
    n.rep <- min(n.starts, ncol(spoints.usr))
    spoints[, (n.starts - n.rep + 1):n.starts] <- spoints.usr
  }
  spoints
}

# optimHypers_* (R/3_training_SF.R:122-223): stats::optim stays in charge (same L-BFGS-B, same control);
# fn = one point, gr = R's own central-difference rule evaluated for all 2d neighbours in ONE device call (quirk Q10)
optim_B200 <- function(sp, bnds, prob, kerType, var.known, full_theta, nugget, control.optim) {
  fn <- function(th) negLogLik_funGp_B200(matrix(th, ncol = 1), prob, kerType, var.known, full_theta, nugget)
  gr <- function(th) {
    d <- length(th); eps <- 1e-3
    P <- matrix(th, d, 2 * d)
    for (i in seq_len(d)) { P[i, 2 * i - 1] <- min(th[i] + eps, bnds[2, i]); P[i, 2 * i] <- max(th[i] - eps, bnds[1, i]) }
    v <- negLogLik_funGp_B200(P, prob, kerType, var.known, full_theta, nugget)
    (v[seq(1, 2 * d, 2)] - v[seq(2, 2 * d, 2)]) / (P[cbind(seq_len(d), seq(1, 2 * d, 2))] - P[cbind(seq_len(d), seq(2, 2 * d, 2))])
  }
  optim(par = sp, fn = fn, gr = gr, method = "L-BFGS-B", lower = bnds[1, ], upper = bnds[2, ], control = control.optim)
}

# setHypers_S / _F / _SF (R/3_training_SF.R:4-49, R/3_training_S.R, R/3_training_F.R): one body for the three input
# cases.  ds / dfl = number of scalar / functional length-scales of the problem; the known ones are spliced into every
# point by full_theta (the three cases of R/3_training_SF.R:235-244).  Same return value: list(hypers, convg, nllik).
setHypers_B200 <- function(prob, ds, dfl, kerType, var.known, ls_s.known, ls_f.known, n.starts, n.presample,
                           spoints.usr, nugget, trace, pbars, control.optim) {
  s.free <- ds > 0 && is.null(ls_s.known); f.free <- dfl > 0 && is.null(ls_f.known)
  full_theta <- function(th) {
    th <- as.matrix(th)
    rbind(if (ds > 0) (if (s.free) th[seq_len(ds), , drop = FALSE] else matrix(ls_s.known, ds, ncol(th))),
          if (dfl > 0) (if (f.free) th[(if (s.free) ds else 0) + seq_len(dfl), , drop = FALSE] else matrix(ls_f.known, dfl, ncol(th))))
  }
  if (!s.free && !f.free) {                                # all length-scales known: variance only (:7-18)
    if (trace) message("** Computing optimal variance...")
    r <- .Call(fgpR_nll_batch, prob, .KER[[kerType]], nugget, full_theta(matrix(0, 0, 1)), NULL)
    if (r[[3]][1] > 0L) stop(sprintf("the leading minor of order %d is not positive definite", r[[3]][1]))
    return(list(hypers = c(r[[2]][1], ls_s.known, ls_f.known), convg = as.numeric(NA), nllik = as.numeric(NA)))
  }
  free <- c(if (s.free) seq_len(ds), if (f.free) ds + seq_len(dfl))
  bnds <- setBounds_B200(prob, free)
  if (trace) message("** Presampling...")
  spoints <- setSPoints_B200(bnds, prob, kerType, var.known, full_theta, n.starts, n.presample, spoints.usr, nugget)
  if (trace) message("** Optimising hyperparameters...")
  if (n.starts == 1) {
    optOut <- optim_B200(as.numeric(spoints), bnds, prob, kerType, var.known, full_theta, nugget, control.optim)
  } else {                                                 # sequential multistart with tryCatch per start (:140-158)
    optOutList <- list()
    if (pbars) pb <- txtProgressBar(min = 0, max = n.starts, style = 3)
    for (i in 1:n.starts) {
      modeval <- tryCatch(optim_B200(as.numeric(spoints[, i]), bnds, prob, kerType, var.known, full_theta, nugget, control.optim),
                          error = function(e) e)
      if (!inherits(modeval, "error")) optOutList[[length(optOutList) + 1]] <- modeval
      if (pbars) setTxtProgressBar(pb, i)
    }
    if (pbars) close(pb)
    if (length(optOutList) == 0) stop("All model optimizations crashed!")
    optOut <- optOutList[[which.min(sapply(optOutList, function(sol) sol$value))]]
  }
  if (trace) message("** Hyperparameters done!")
  theta <- as.numeric(full_theta(matrix(optOut$par, ncol = 1)))
  sig2 <- if (is.null(var.known)) .Call(fgpR_nll_batch, prob, .KER[[kerType]], nugget, matrix(theta, ncol = 1), NULL)[[2]][1] else var.known
  list(hypers = c(sig2, theta), convg = optOut$convergence, nllik = optOut$value)
}

# preMats_* (R/4_prediction_SF.R:24-32); the handle rides on the returned list
preMats_B200 <- function(prob, sig2, thetas, kerType, nugget) {
  r <- .Call(fgpR_premats, prob, .KER[[kerType]], nugget, sig2, as.numeric(thetas))
  structure(list(L = r[[1]], LInvY = r[[2]]), b200prob = prob)
}

# update() fast paths (R/6_updating.R): the refit with retained hyper-parameters re-uses the old model's factor
preMats_reuse_B200 <- function(prob.new, model.old) {
  r <- .Call(fgpR_premats_reuse, prob.new, b200_handle(model.old))
  structure(list(L = r[[1]], LInvY = r[[2]]), b200prob = prob.new)
}
preMats_newvar_B200 <- function(model, var.new) {          # upd_subHypers with var.sb only (:510-524)
  r <- .Call(fgpR_update_var, b200_handle(model), var.new)
  structure(list(L = r[[1]], LInvY = r[[2]]), b200prob = b200_handle(model))
}

# makePreds_* (R/4_prediction_SF.R:2-21); new curves are projected first (R/1_fgpm_Class.R:844-845)
makePreds_B200 <- function(prob, sIn.pr, coefs.pr, n.pr, detail) {
  r <- .Call(fgpR_predict, prob, as.integer(n.pr), sIn.pr, coefs.pr, as.integer(detail == "full"))
  out <- list(mean = r[[1]], sd = r[[2]])
  if (detail == "full") { out$K.tp <- r[[3]]; out$K.pp <- r[[4]] }
  out
}

# makeSims_* (R/5_simulation_SF.R:3-23): set.seed + rnorm stay in R (:13-14) so the stream is the reference's
makeSims_B200 <- function(prob, sIn.sm, coefs.sm, n.sm, nsim, nugget.sm, detail, seed) {
  if (!is.null(seed)) set.seed(seed)
  Z <- matrix(rnorm(n.sm * nsim), n.sm, nsim)
  r <- .Call(fgpR_simulate, prob, as.integer(n.sm), sIn.sm, coefs.sm, Z, nugget.sm)
  if (detail == "light") list(sims = r[[1]]) else list(sims = r[[1]], mean = r[[2]], sd = r[[3]])
}

# getOut_loocv (R/8_outilsStats.R:1-7): same return value (n x 1 matrix of leave-one-out predictions)
# (the device keeps the result with the model, so getFitness followed by plot(model) computes it once)
getOut_loocv <- function(model) .Call(fgpR_loocv, b200_handle(model))[[1]]

# eval_loocv_ACO / eval_houtv_ACO (R/3_ant_admin.R:592-648, 662-802): one device call per generation.  The presample
# uniforms are drawn HERE, in ant (and replicate) order, exactly as the sequential loop over fitNtests_ACO -> fgpm ->
# setSPoints_* would draw them (runif(n.ls * n.presample) per fit, R/3_training_SF.R:91), so R's RNG stream -- and with
# it the ants the colony builds next -- is unchanged.  n.ls per ant comes from its structure (getOwners).
eval_ACO_B200 <- function(sIn, fIn, sOut, extargs, ants, n.ls, ind.vl = NULL) {
  n.rep <- if (is.null(ind.vl)) 1L else ncol(ind.vl)
  n.pre <- max(extargs$n.presample, extargs$n.starts)
  sizes <- rep(n.ls * n.pre, each = n.rep)
  u01 <- runif(sum(sizes))
  antsT <- t(ants); storage.mode(antsT) <- "integer"            # the colony builds `ants` as a double matrix
  if (!is.null(ind.vl)) { ind.vl <- as.matrix(ind.vl); storage.mode(ind.vl) <- "integer" }
  r <- .Call(fgpR_fit_batch, b200_ctx(), sIn, fIn, as.numeric(sOut), antsT, extargs$nugget, as.integer(extargs$n.starts),
             as.integer(extargs$n.presample), u01, as.numeric(c(0, cumsum(sizes))[seq_along(sizes)]), as.integer(max(n.ls)), ind.vl)
  fitness <- matrix(r[[1]], nrow = n.rep)                       # NaN = crashed fit (logged as a crash, fitness NA)
  list(fitness = apply(fitness, 2, function(v) if (all(is.nan(v))) NA else mean(v[!is.nan(v)])), nll = r[[2]], sig2 = r[[3]],
       hypers = r[[4]], info = r[[5]])
}

# eval_loocv_ACO / eval_houtv_ACO keep their signatures and their return value -- a list with one entry
# list(args, model, fitness, ant) per ant (R/3_ant_admin.R:592-648, 662-802) -- but the `model` entry is a promise:
# the hyper-parameters the device found.  run_ACO only ever uses the model of the best ant (R/3_ant_search.R:138-149);
# b200_materialize turns that one into a real fgpm object with a known-hyper-parameter fgpm() call (one preMats).
eval_ACO_as_list_B200 <- function(sIn, fIn, sOut, extargs, base, ants, ind.vl = NULL) {
  args <- lapply(seq_len(nrow(ants)), function(i) formatSol_ACO(ants[i, ], sIn, fIn, base))
  n.ls <- sapply(args, function(a) (if (is.null(a$sIn)) 0 else ncol(a$sIn)) +
                   sum(mapply(function(dt, p, f) if (dt == "L2_bygroup") 1 else if (p > 0) p else ncol(f), a$f_disType, a$f_pdims, a$fIn)))
  r <- eval_ACO_B200(sIn, fIn, sOut, extargs, ants, n.ls, ind.vl)
  lapply(seq_len(nrow(ants)), function(i) {
    n.rep <- if (is.null(ind.vl)) 1L else ncol(ind.vl)
    it <- (i - 1L) * n.rep + 1L                                   # hold-out: hyper-parameters of the first replicate
    promise <- list(args = args[[i]], sig2 = r$sig2[it], theta = r$hypers[seq_len(n.ls[i]), it], extargs = extargs, sOut = sOut)
    list(args[[i]][-(1:2)], structure(promise, class = "b200_model_promise"), r$fitness[i], ants[i, ])
  })
}
b200_materialize <- function(m) {
  if (!inherits(m, "b200_model_promise")) return(m)
  a <- m$args
  ds <- if (is.null(a$sIn)) 0 else ncol(a$sIn)
  fgpm(sIn = a$sIn, fIn = a$fIn, sOut = m$sOut, kerType = a$kerType, f_disType = a$f_disType, f_pdims = a$f_pdims,
       f_basType = a$f_basType, var.hyp = m$sig2, ls_s.hyp = if (ds > 0) m$theta[seq_len(ds)] else NULL,
       ls_f.hyp = if (length(m$theta) > ds) m$theta[-seq_len(ds)] else NULL, nugget = m$extargs$nugget, trace = FALSE, pbars = FALSE)
}

# Same names and signatures as the reference (R/3_ant_admin.R:592, 662): collated last, these definitions replace the
# sequential loops over fitNtests_ACO.  One batched device call per generation; par.clust / time.lim are not needed
# (the whole generation returns in about the time one ant took before).
eval_loocv_ACO <- function(sIn, fIn, sOut, extargs, base, ants, time.str, time.lim, pbars, par.clust)
  eval_ACO_as_list_B200(sIn, fIn, sOut, extargs, base, ants)
eval_houtv_ACO <- function(sIn, fIn, sOut, extargs, base, ants, ind.vl, time.str, time.lim, pbars, par.clust)
  eval_ACO_as_list_B200(sIn, fIn, sOut, extargs, base, ants, ind.vl)
