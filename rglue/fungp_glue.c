/*
 * fungp_glue.c -- the thin R .Call layer over libfungp_b200.so (include/fungp_b200.h).
 *
 * This file belongs in the R package's new src/ (next to init.c and Makevars from this directory).  It contains NO
 * numerics: every wrapper unpacks SEXPs (column-major REALSXP exactly as the C ABI expects), calls one fgp_* entry
 * point and packs the result.  R is not installed in the build image, so this file is syntax-checked against the
 * minimal declarations in rglue/mock/ (tests/test_rglue_syntax.py) and has never been linked against libR.
 *
 * Ownership: inputs are read-only; outputs are allocVector/allocMatrix + PROTECT; device handles are R
 * external pointers with C finalizers; an info > 0 from the library becomes base::chol's own error message; Rf_error
 * long-jumps, so nothing is left allocated on the C side when it is raised (the library owns all temporaries).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include "fungp_b200.h"

static void ctx_finalizer(SEXP p) {
    fgp_ctx* c = (fgp_ctx*)R_ExternalPtrAddr(p);
    if (c) { fgp_ctx_destroy(c); R_ClearExternalPtr(p); }
}
static void prob_finalizer(SEXP p) {
    fgp_problem* q = (fgp_problem*)R_ExternalPtrAddr(p);
    if (q) { fgp_problem_destroy(q); R_ClearExternalPtr(p); }
}
static fgp_ctx* get_ctx(SEXP p) {
    fgp_ctx* c = (fgp_ctx*)R_ExternalPtrAddr(p);
    if (!c) Rf_error("funGp: the GPU context is gone (was the session restored from disk?)");
    return c;
}
static fgp_problem* get_prob(SEXP p) {
    fgp_problem* q = (fgp_problem*)R_ExternalPtrAddr(p);
    if (!q) Rf_error("funGp: the device problem handle is NULL; rebuild it from the model's slots");
    return q;
}
static void check(int rc, fgp_ctx* ctx, const char* what) {
    if (rc > 0) Rf_error("the leading minor of order %d is not positive definite", rc);   /* base::chol's message */
    if (rc < 0) Rf_error("funGp/%s failed (%d): %s", what, rc, ctx ? fgp_last_error(ctx) : "no context");
}
/* the context a problem handle was created in (kept alive as the external pointer's protected value): the library's
 * error text for a failed call on that problem */
static fgp_ctx* prob_ctx(SEXP prob) {
    SEXP c = R_ExternalPtrProtected(prob);
    return (TYPEOF(c) == EXTPTRSXP) ? (fgp_ctx*)R_ExternalPtrAddr(c) : NULL;
}

/* is this external pointer still alive?  FALSE after save()/load() (nil address) or for anything that is not a pointer */
SEXP fgpR_ptr_valid(SEXP p) {
    return Rf_ScalarLogical(TYPEOF(p) == EXTPTRSXP && R_ExternalPtrAddr(p) != NULL);
}

/* .Call(C_fgpR_ctx_create, n_gpus)  -- options(funGp.gpus=) / FUNGP_GPUS, 0 = all visible */
SEXP fgpR_ctx_create(SEXP n_gpus) {
    fgp_ctx* c = NULL;
    int rc = fgp_ctx_create(Rf_asInteger(n_gpus), NULL, &c);
    if (rc != 0) Rf_error("funGp: no sm_100 CUDA device available (code %d); this build has no CPU path", rc);
    SEXP p = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* dimReduction for one functional input (R/7_dimRedFunctions.R:13-27) or re-projection with a stored basis
 * (R/1_fgpm_Class.R:844-845): .Call(C_fgpR_project, ctx, f, p, basis_type, B_or_NULL) -> list(basis, J, coefs) */
SEXP fgpR_project(SEXP ctx, SEXP f, SEXP p, SEXP basis_type, SEXP B) {
    const int n = Rf_nrows(f), k = Rf_ncols(f), pp = Rf_asInteger(p), pe = pp ? pp : k;
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP Bo = PROTECT(Rf_allocMatrix(REALSXP, k, pe)), J = PROTECT(Rf_allocMatrix(REALSXP, pe, pe));
    SEXP co = PROTECT(Rf_allocMatrix(REALSXP, n, pe));
    int rc = fgp_project(get_ctx(ctx), REAL(f), n, k, pp, Rf_asInteger(basis_type), Rf_isNull(B) ? NULL : REAL(B), REAL(Bo),
                         REAL(J), REAL(co));
    check(rc > 0 ? -1 : rc, get_ctx(ctx), "project");
    SET_VECTOR_ELT(out, 0, Bo); SET_VECTOR_ELT(out, 1, J); SET_VECTOR_ELT(out, 2, co);
    UNPROTECT(4);
    return out;
}

/* what setDistMatrix_S / setDistMatrix_F receive (R/7_distanceFunctions.R:3-22), kept on the device instead of the
 * materialised sMs / fMs lists: .Call(C_fgpR_problem_create, ctx, sIn_or_NULL, coefs(list), J(list), disType(int), sOut) */
SEXP fgpR_problem_create(SEXP ctx, SEXP sIn, SEXP coefs, SEXP J, SEXP dis, SEXP sOut) {
    const int n = Rf_length(sOut), ds = Rf_isNull(sIn) ? 0 : Rf_ncols(sIn), df = Rf_length(coefs);
    const double* cp[64]; const double* jp[64]; int p[64];
    if (df > 64) Rf_error("funGp: more than 64 functional inputs");
    for (int i = 0; i < df; ++i) {
        cp[i] = REAL(VECTOR_ELT(coefs, i)); jp[i] = REAL(VECTOR_ELT(J, i)); p[i] = Rf_ncols(VECTOR_ELT(coefs, i));
    }
    fgp_problem* q = NULL;
    int rc = fgp_problem_create(get_ctx(ctx), n, ds, ds ? REAL(sIn) : NULL, df, p, INTEGER(dis), cp, jp, REAL(sOut), &q);
    check(rc, get_ctx(ctx), "problem_create");
    SEXP e = PROTECT(R_MakeExternalPtr(q, R_NilValue, ctx));   /* prot = ctx keeps the context alive */
    R_RegisterCFinalizerEx(e, prob_finalizer, TRUE);
    UNPROTECT(1);
    return e;
}

/* setBounds_* : sapply(Ms, max)  (R/3_training_SF.R:57-74) */
SEXP fgpR_dist_max(SEXP prob) {
    fgp_problem* q = get_prob(prob);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, fgp_problem_nls(q)));
    check(fgp_dist_max(q, REAL(out)), prob_ctx(prob), "dist_max");
    UNPROTECT(1);
    return out;
}

/* one element of c(sMs, fMs) on demand (API parity for code that inspects the distance lists) */
SEXP fgpR_dist_matrix(SEXP prob, SEXP l) {
    fgp_problem* q = get_prob(prob);
    const int n = fgp_problem_n(q);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, n));
    check(fgp_dist_materialize(q, Rf_asInteger(l) - 1, REAL(out)), prob_ctx(prob), "dist_materialize");
    UNPROTECT(1);
    return out;
}

/* setR(...) * setR(...) + diag(nugget)  (R/7_correlFunctions.R:1-15, R/3_training_SF.R:248) */
SEXP fgpR_assemble(SEXP prob, SEXP ker, SEXP nugget, SEXP theta) {
    fgp_problem* q = get_prob(prob);
    const int n = fgp_problem_n(q);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, n));
    check(fgp_assemble(q, Rf_asInteger(ker), Rf_asReal(nugget), REAL(theta), REAL(out)), prob_ctx(prob), "assemble");
    UNPROTECT(1);
    return out;
}

/* negLogLik_funGp_* for the columns of `thetas` (d x B) in one launch sequence (R/3_training_SF.R:231-258);
 * var_known = NULL or a length-1 numeric.  -> list(nll, sig2, info); the caller decides error vs skip (quirk Q6) */
SEXP fgpR_nll_batch(SEXP prob, SEXP ker, SEXP nugget, SEXP thetas, SEXP var_known) {
    fgp_problem* q = get_prob(prob);
    const int B = Rf_isMatrix(thetas) ? Rf_ncols(thetas) : 1;
    double vk = Rf_isNull(var_known) ? 0.0 : Rf_asReal(var_known);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP nll = PROTECT(Rf_allocVector(REALSXP, B)), s2 = PROTECT(Rf_allocVector(REALSXP, B));
    SEXP info = PROTECT(Rf_allocVector(INTSXP, B));
    int rc = fgp_nll_batch(q, Rf_asInteger(ker), Rf_asReal(nugget), B, REAL(thetas), Rf_isNull(var_known) ? NULL : &vk, REAL(nll),
                           REAL(s2), INTEGER(info));
    check(rc, prob_ctx(prob), "nll_batch");
    SET_VECTOR_ELT(out, 0, nll); SET_VECTOR_ELT(out, 1, s2); SET_VECTOR_ELT(out, 2, info);
    UNPROTECT(4);
    return out;
}

/* preMats_* (R/4_prediction_SF.R:24-32) -> list(L, LInvY); the factor also stays on the device for predict/simulate */
SEXP fgpR_premats(SEXP prob, SEXP ker, SEXP nugget, SEXP sig2, SEXP theta) {
    fgp_problem* q = get_prob(prob);
    const int n = fgp_problem_n(q);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP L = PROTECT(Rf_allocMatrix(REALSXP, n, n)), z = PROTECT(Rf_allocMatrix(REALSXP, n, 1));
    check(fgp_premats(q, Rf_asInteger(ker), Rf_asReal(nugget), Rf_asReal(sig2), REAL(theta), REAL(L), REAL(z)), prob_ctx(prob), "premats");
    SET_VECTOR_ELT(out, 0, L); SET_VECTOR_ELT(out, 1, z);
    UNPROTECT(3);
    return out;
}

static int new_points(SEXP coefsNew, const double** cp) {
    const int df = Rf_length(coefsNew);
    if (df > 64) Rf_error("funGp: more than 64 functional inputs");
    for (int i = 0; i < df; ++i) cp[i] = REAL(VECTOR_ELT(coefsNew, i));
    return df;
}

/* makePreds_* (R/4_prediction_SF.R:2-21) -> list(mean, sd [, K.tp, K.pp]) */
SEXP fgpR_predict(SEXP prob, SEXP m_, SEXP sNew, SEXP coefsNew, SEXP detail) {
    fgp_problem* q = get_prob(prob);
    const int n = fgp_problem_n(q), m = Rf_asInteger(m_), full = Rf_asInteger(detail);
    const double* cp[64];
    new_points(coefsNew, cp);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, full ? 4 : 2));
    SEXP mean = PROTECT(Rf_allocVector(REALSXP, m)), sd = PROTECT(Rf_allocVector(REALSXP, m));
    SEXP Ktp = R_NilValue, Kpp = R_NilValue;
    if (full) { Ktp = PROTECT(Rf_allocMatrix(REALSXP, n, m)); Kpp = PROTECT(Rf_allocMatrix(REALSXP, m, m)); }
    int rc = fgp_predict(q, m, Rf_isNull(sNew) ? NULL : REAL(sNew), cp, full, REAL(mean), REAL(sd), full ? REAL(Ktp) : NULL,
                         full ? REAL(Kpp) : NULL);
    check(rc, prob_ctx(prob), "predict");
    SET_VECTOR_ELT(out, 0, mean); SET_VECTOR_ELT(out, 1, sd);
    if (full) { SET_VECTOR_ELT(out, 2, Ktp); SET_VECTOR_ELT(out, 3, Kpp); }
    UNPROTECT(full ? 5 : 3);
    return out;
}

/* makeSims_* (R/5_simulation_SF.R:3-23); Z = matrix(rnorm(n.sm*nsim), n.sm, nsim) is drawn in R after set.seed(seed) */
SEXP fgpR_simulate(SEXP prob, SEXP m_, SEXP sNew, SEXP coefsNew, SEXP Z, SEXP nugget_sm) {
    fgp_problem* q = get_prob(prob);
    const int m = Rf_asInteger(m_), nsim = Rf_ncols(Z);
    const double* cp[64];
    new_points(coefsNew, cp);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP sims = PROTECT(Rf_allocMatrix(REALSXP, nsim, m)), mean = PROTECT(Rf_allocVector(REALSXP, m));
    SEXP sd = PROTECT(Rf_allocVector(REALSXP, m));
    int rc = fgp_simulate(q, m, Rf_isNull(sNew) ? NULL : REAL(sNew), cp, nsim, REAL(Z), Rf_asReal(nugget_sm), REAL(sims), REAL(mean),
                          REAL(sd));
    check(rc, prob_ctx(prob), "simulate");
    SET_VECTOR_ELT(out, 0, sims); SET_VECTOR_ELT(out, 1, mean); SET_VECTOR_ELT(out, 2, sd);
    UNPROTECT(4);
    return out;
}

/* getOut_loocv (R/8_outilsStats.R:1-7) + Q2 (R/1_Xfgpm_Class.R:541) -> list(y_pre, q2) */
SEXP fgpR_loocv(SEXP prob) {
    fgp_problem* q = get_prob(prob);
    const int n = fgp_problem_n(q);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP yp = PROTECT(Rf_allocMatrix(REALSXP, n, 1)), q2 = PROTECT(Rf_allocVector(REALSXP, 1));
    check(fgp_loocv(q, REAL(yp), REAL(q2)), prob_ctx(prob), "loocv");
    SET_VECTOR_ELT(out, 0, yp); SET_VECTOR_ELT(out, 1, q2);
    UNPROTECT(3);
    return out;
}

/* output substitution of update() (R/6_updating.R:154-160) */
SEXP fgpR_set_y(SEXP prob, SEXP y) {
    check(fgp_problem_set_y(get_prob(prob), REAL(y)), prob_ctx(prob), "set_y");
    return R_NilValue;
}

/* eval_loocv_ACO / eval_houtv_ACO for one generation (R/3_ant_admin.R:592-648, 662-802):
 * .Call(fgpR_fit_batch, ctx, sIn_or_NULL, fIn(list of n x k_j matrices), sOut, t(ants) (integer, one candidate per COLUMN so
 *       that each candidate is contiguous), nugget, n.starts, n.presample, u01 (the runif draws, concatenated per work item),
 *       u01_off (numeric offsets, one per work item), dmax, ind.vl_or_NULL (integer n.vl x n.rep))
 * -> list(fitness, nll, sig2, theta (dmax x items), info); items = ants, or (ant, replicate) pairs ant-major. */
SEXP fgpR_fit_batch(SEXP ctx, SEXP sIn, SEXP fIn, SEXP sOut, SEXP antsT, SEXP nugget, SEXP n_starts, SEXP n_presample, SEXP u01,
                    SEXP u01_off, SEXP dmax_, SEXP ind_vl) {
    const int n = Rf_length(sOut), ds = Rf_isNull(sIn) ? 0 : Rf_ncols(sIn), df = Rf_length(fIn);
    const int n_cand = Rf_ncols(antsT), dmax = Rf_asInteger(dmax_);
    const int n_rep = Rf_isNull(ind_vl) ? 0 : Rf_ncols(ind_vl), n_vl = Rf_isNull(ind_vl) ? 0 : Rf_nrows(ind_vl);
    const int items = n_cand * (n_rep > 0 ? n_rep : 1);
    const double* fp[64]; int k[64];
    if (df > 64) Rf_error("funGp: more than 64 functional inputs");
    if (Rf_length(u01_off) != items) Rf_error("funGp: u01_off must have one entry per work item");
    long long* off = (long long*)R_alloc((size_t)(items > 0 ? items : 1), sizeof(long long));
    for (int j = 0; j < df; ++j) { fp[j] = REAL(VECTOR_ELT(fIn, j)); k[j] = Rf_ncols(VECTOR_ELT(fIn, j)); }
    for (int i = 0; i < items; ++i) off[i] = (long long)REAL(u01_off)[i];
    /* the colony's ant matrix and ind.vl are double matrices in R unless coerced: never read them through INTEGER() as is */
    antsT = PROTECT(Rf_coerceVector(antsT, INTSXP));
    ind_vl = PROTECT(Rf_isNull(ind_vl) ? ind_vl : Rf_coerceVector(ind_vl, INTSXP));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
    SEXP fit = PROTECT(Rf_allocVector(REALSXP, items)), nll = PROTECT(Rf_allocVector(REALSXP, items));
    SEXP s2 = PROTECT(Rf_allocVector(REALSXP, items)), th = PROTECT(Rf_allocMatrix(REALSXP, dmax, items));
    SEXP info = PROTECT(Rf_allocVector(INTSXP, items));
    int rc;
    if (n_rep == 0)
        rc = fgp_fit_batch(get_ctx(ctx), n, ds, ds ? REAL(sIn) : NULL, df, k, fp, REAL(sOut), n_cand, INTEGER(antsT), Rf_asReal(nugget),
                           Rf_asInteger(n_starts), Rf_asInteger(n_presample), REAL(u01), off, dmax, REAL(fit), REAL(nll), REAL(s2),
                           REAL(th), INTEGER(info));
    else
        rc = fgp_fit_batch_holdout(get_ctx(ctx), n, ds, ds ? REAL(sIn) : NULL, df, k, fp, REAL(sOut), n_cand, INTEGER(antsT),
                                   Rf_asReal(nugget), Rf_asInteger(n_starts), Rf_asInteger(n_presample), REAL(u01), off, dmax, n_rep,
                                   n_vl, INTEGER(ind_vl), REAL(fit), REAL(nll), REAL(s2), REAL(th), INTEGER(info));
    check(rc, get_ctx(ctx), "fit_batch");
    SET_VECTOR_ELT(out, 0, fit); SET_VECTOR_ELT(out, 1, nll); SET_VECTOR_ELT(out, 2, s2); SET_VECTOR_ELT(out, 3, th);
    SET_VECTOR_ELT(out, 4, info);
    UNPROTECT(8);
    return out;
}

/* (de)serialisation: hand the stored preMats of a loaded model back to a fresh problem handle -- no factorisation
 * .Call(fgpR_restore, prob, ker, nugget, sig2, theta, L, LInvY) */
SEXP fgpR_restore(SEXP prob, SEXP ker, SEXP nugget, SEXP sig2, SEXP theta, SEXP L, SEXP z) {
    check(fgp_problem_restore(get_prob(prob), Rf_asInteger(ker), Rf_asReal(nugget), Rf_asReal(sig2), REAL(theta), REAL(L), REAL(z)),
          prob_ctx(prob), "problem_restore");
    return R_NilValue;
}

/* update() fast paths (R/6_updating.R): -> list(L, LInvY) of the updated model */
SEXP fgpR_update_var(SEXP prob, SEXP sig2_new) {
    fgp_problem* q = get_prob(prob);
    const int n = fgp_problem_n(q);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP L = PROTECT(Rf_allocMatrix(REALSXP, n, n)), z = PROTECT(Rf_allocMatrix(REALSXP, n, 1));
    check(fgp_update_var(q, Rf_asReal(sig2_new), REAL(L), REAL(z)), prob_ctx(prob), "update_var");
    SET_VECTOR_ELT(out, 0, L); SET_VECTOR_ELT(out, 1, z);
    UNPROTECT(3);
    return out;
}
SEXP fgpR_update_y(SEXP prob, SEXP y) {
    fgp_problem* q = get_prob(prob);
    SEXP z = PROTECT(Rf_allocMatrix(REALSXP, fgp_problem_n(q), 1));
    check(fgp_update_y(q, REAL(y), REAL(z)), prob_ctx(prob), "update_y");
    UNPROTECT(1);
    return z;
}
SEXP fgpR_premats_reuse(SEXP prob_new, SEXP prob_old) {
    fgp_problem* q = get_prob(prob_new);
    const int n = fgp_problem_n(q);
    int reused = 0;
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP L = PROTECT(Rf_allocMatrix(REALSXP, n, n)), z = PROTECT(Rf_allocMatrix(REALSXP, n, 1));
    check(fgp_premats_reuse(q, get_prob(prob_old), REAL(L), REAL(z), &reused), prob_ctx(prob_new), "premats_reuse");
    SET_VECTOR_ELT(out, 0, L); SET_VECTOR_ELT(out, 1, z); SET_VECTOR_ELT(out, 2, Rf_ScalarInteger(reused));
    UNPROTECT(3);
    return out;
}
