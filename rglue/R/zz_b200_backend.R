# zz_b200_backend.R -- re-points the INTERNAL hot-path functions of funGp at libfungp_b200 through .Call.
# Collated last (DESCRIPTION Collate), so these definitions replace the R bodies of the same names; the exported API
# (fgpm, predict, simulate, update, fgpm_factory) and every signature visible to users is unchanged.
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

# preMats_* (R/4_prediction_SF.R:24-32)
preMats_B200 <- function(prob, sig2, thetas, kerType, nugget) {
  r <- .Call(fgpR_premats, prob, .KER[[kerType]], nugget, sig2, as.numeric(thetas))
  list(L = r[[1]], LInvY = r[[2]])
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
getOut_loocv <- function(model) .Call(fgpR_loocv, model@b200prob)[[1]]

# eval_loocv_ACO / eval_houtv_ACO (R/3_ant_admin.R:592-648, 662-802): one device call per generation.  The presample
# uniforms are drawn HERE, in ant (and replicate) order, exactly as the sequential loop over fitNtests_ACO -> fgpm ->
# setSPoints_* would draw them (runif(n.ls * n.presample) per fit, R/3_training_SF.R:91), so R's RNG stream -- and with
# it the ants the colony builds next -- is unchanged.  n.ls per ant comes from its structure (getOwners).
eval_ACO_B200 <- function(sIn, fIn, sOut, extargs, ants, n.ls, ind.vl = NULL) {
  n.rep <- if (is.null(ind.vl)) 1L else ncol(ind.vl)
  n.pre <- max(extargs$n.presample, extargs$n.starts)
  sizes <- rep(n.ls * n.pre, each = n.rep)
  u01 <- runif(sum(sizes))
  r <- .Call(fgpR_fit_batch, b200_ctx(), sIn, fIn, as.numeric(sOut), t(ants), extargs$nugget, as.integer(extargs$n.starts),
             as.integer(extargs$n.presample), u01, as.numeric(c(0, cumsum(sizes))[seq_along(sizes)]), as.integer(max(n.ls)), ind.vl)
  fitness <- matrix(r[[1]], nrow = n.rep)                       # NaN = crashed fit (logged as a crash, fitness NA)
  list(fitness = apply(fitness, 2, function(v) if (all(is.nan(v))) NA else mean(v[!is.nan(v)])), nll = r[[2]], hypers = r[[4]])
}
