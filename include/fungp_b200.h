/*
 * fungp_b200.h -- plain-C ABI of libfungp_b200.so: the sm_100a (B200) implementation of funGp's
 * Gaussian-process hot path.  Every entry point replaces the body of an *internal R function* of the
 * reference package funGp 1.0.0 (the reference has no FFI of its own: no src/, no useDynLib --
 * NAMESPACE:1-78), so each declaration cites the reference file:line whose arithmetic it takes over.
 * The R-side .Call stubs that bind these symbols are in rglue/ and described in INTEGRATION.md.
 *
 * Conventions
 *   - all matrices are column-major doubles exactly as R's REALSXP + dim (n x k, leading dimension n);
 *   - every pointer is a HOST pointer; the library never keeps a host pointer after the call returns
 *     and never writes through an input pointer;
 *   - return value: 0 = ok; >0 = LAPACK-style info of base::chol ("the leading minor of order <info> is
 *     not positive definite", what reaches users today through R/3_training_SF.R:249); <0 = bad argument
 *     or CUDA failure, text in fgp_last_error();  batched calls report per-element info[] instead and
 *     return 0 unless the call itself failed;
 *   - integer codes follow the reference's own level order (R/1_Xfgpm_Class.R:457-459,
 *     R/3_ant_admin.R:301-373):  ker 1 = gauss, 2 = matern5_2, 3 = matern3_2;
 *     dis_type 1 = L2_bygroup, 2 = L2_byindex;  basis_type 1 = B-splines, 2 = PCA;  detail 0 = light, 1 = full;
 *   - there is no CPU fallback: without a CUDA device fgp_ctx_create() fails.
 */
#ifndef FUNGP_B200_H
#define FUNGP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fgp_ctx fgp_ctx;         /* devices, streams, workspaces                                   */
typedef struct fgp_problem fgp_problem; /* one data set (coordinates + output) resident on the device(s)  */

enum { FGP_KER_GAUSS = 1, FGP_KER_MATERN52 = 2, FGP_KER_MATERN32 = 3 };
enum { FGP_DIST_BYGROUP = 1, FGP_DIST_BYINDEX = 2 };
enum { FGP_BASIS_BSPLINES = 1, FGP_BASIS_PCA = 2 };
enum { FGP_DETAIL_LIGHT = 0, FGP_DETAIL_FULL = 1 };

/* error codes (<0) */
enum { FGP_ERR_ARG = -1, FGP_ERR_CUDA = -2, FGP_ERR_NODEVICE = -3, FGP_ERR_STATE = -4, FGP_ERR_ALLOC = -5 };

/* ---- context ------------------------------------------------------------------------------------- */
/* n_devices = 0 -> all visible GPUs; device_ids may be NULL (= 0..n_devices-1).  One host worker per GPU;
 * batched calls shard their independent work items over the GPUs, nothing else does. */
int fgp_ctx_create(int n_devices, const int* device_ids, fgp_ctx** out);
void fgp_ctx_destroy(fgp_ctx* ctx);
int fgp_ctx_device_count(const fgp_ctx* ctx);
const char* fgp_last_error(const fgp_ctx* ctx);
const char* fgp_version(void);

/* ---- dimension reduction: R/7_dimRedFunctions.R:7-60 --------------------------------------------- */
/* proj_bsplines, R/7_dimRedFunctions.R:40-51: k x p design matrix on the grid 1..k (host code). */
int fgp_bspline_basis(int k, int p, double* B_out /* k*p */);
/* dimReduction for ONE functional input (R/7_dimRedFunctions.R:13-27) and the re-projection of new curves
 * with a stored basis (R/1_fgpm_Class.R:844-845, 863-864, 1044-1045, 1063-1064):
 *   p = 0            -> B = J = I_k, coefs = f                       (:22-25)
 *   B_in != NULL     -> use B_in (k x p) as the basis (re-projection)
 *   else             -> build the basis: B-splines (:40-51) or PCA = first p eigenvectors of cov(f) (:57-60)
 * J = B'B, coefs = t(solve(J, B' f')).  B_out (k x p'), J_out (p' x p'), coefs_out (n x p'), p' = p ? p : k. */
int fgp_project(fgp_ctx* ctx, const double* f, int n, int k, int p, int basis_type, const double* B_in,
                double* B_out, double* J_out, double* coefs_out);
/* Test support only -- never called by the library or the R glue: the host-side pieces of fgp_project (design matrix,
 * eigenvectors, p x p solve) with the contractions over the curves as plain loops, so that the CPU test gate can check
 * them against the oracle on a machine without a GPU.  fgp_project itself fails without a context. */
int fgp_test_project_host(const double* f, int n, int k, int p, int basis_type, const double* B_in, double* B_out,
                          double* J_out, double* coefs_out);
/* Test support, called by nothing in the product: the claim order of the persistent factorisation kernel
 * (csrc/chol_dag.cu) -- task number -> (tile row, tile column, matrix) for T tile columns, B matrices, claim groups of G
 * columns (G divides T).  tests/test_dag_order_cpu.py checks that every task only depends on tasks with smaller numbers
 * (the kernel's spin-waits are deadlock-free exactly then).  Returns 0, or -1 for arguments out of range. */
int fgp_test_dag_order(int T, int B, int G, int matrixMajor, int task, int* i, int* j, int* b);

/* ---- problem = what setDistMatrix_S / setDistMatrix_F are given (R/7_distanceFunctions.R:3-22) ---- */
/* sIn n x ds (may be NULL when ds = 0); for each functional input i: coefs[i] n x p[i], J[i] p[i] x p[i],
 * dis_type[i]; y = sOut (n).  Length-scale order = reference order: scalars, then per functional input one
 * (L2_bygroup) or p[i] (L2_byindex) entries (getOwners, R/7_distanceFunctions.R:48-58).
 * The distance lists sMs/fMs are never materialised: kernels work from the coordinates. */
int fgp_problem_create(fgp_ctx* ctx, int n, int ds, const double* sIn, int df, const int* p, const int* dis_type,
                       const double* const* coefs, const double* const* J, const double* y, fgp_problem** out);
void fgp_problem_destroy(fgp_problem* prob);
int fgp_problem_n(const fgp_problem* prob);
int fgp_problem_nls(const fgp_problem* prob); /* d = number of length-scales */
/* replace sOut (update(): output substitution, R/6_updating.R:154-160) */
int fgp_problem_set_y(fgp_problem* prob, const double* y);

/* setBounds_*: max(Ms[[l]]) for every length-scale (R/3_training_SF.R:57-74, _S.R:50-60, _F.R:50-60). */
int fgp_dist_max(fgp_problem* prob, double* max_out /* d */);
/* API-parity escape hatch: the l-th (0-based) n x n distance matrix of c(sMs, fMs)
 * (L2_byindex R/7_distanceFunctions.R:28-34, L2_bygroup :38-46). */
int fgp_dist_materialize(fgp_problem* prob, int l, double* out /* n*n */);

/* setR(...) [* setR(...)] + nugget*I  (R/7_correlFunctions.R:1-80, R/3_training_SF.R:248); full symmetric n x n. */
int fgp_assemble(fgp_problem* prob, int ker, double nugget, const double* theta /* d */, double* R_out /* n*n */);

/* negLogLik_funGp_{S,F,SF} for B length-scale vectors in one call (R/3_training_SF.R:231-258, _S.R:200-213,
 * _F.R:197-210) with analyticVar_llik (R/3_training_SF.R:266-269).  thetas is d x B (column b = point b).
 * var_known: NULL -> sigma^2 estimated (analyticVar_llik); else the fixed variance (quirk Q1: the formula
 * keeps '+ n').  Outputs nll[B], sig2[B], info[B] (0 or the failing minor's order; nll = NaN there).
 * The B points are independent: they are sharded over the context's GPUs. */
int fgp_nll_batch(fgp_problem* prob, int ker, double nugget, int B, const double* thetas, const double* var_known,
                  double* nll, double* sig2, int* info);

/* preMats_*: K = sig2*(R + nugget I), L = t(chol(K)), LInvY = L^-1 y  (R/4_prediction_SF.R:24-32).
 * L_out (n x n, explicit zeros above the diagonal) and LInvY_out (n) may be NULL; the factor always stays
 * resident on device 0 inside the problem for predict / simulate / loocv, with `predict_reserve` (option, default 2048)
 * free rows under it in which predict / simulate solve new points in place.  Large outputs (L_out here, K.tp / K.pp /
 * sims below) travel in column chunks through pinned double buffers; L_out moves its lower triangle only. */
int fgp_premats(fgp_problem* prob, int ker, double nugget, double sig2, const double* theta, double* L_out,
                double* LInvY_out);

/* (De)serialisation.  A saved fgpm object keeps preMats$L and preMats$LInvY in its slots but not the
 * device handle (external pointers do not survive save()/load()): after load() the glue re-creates the problem from the
 * slots (fgp_problem_create) and hands the stored factor back instead of assembling and factoring again.
 * L is n x n column-major with zeros above the diagonal, exactly what fgp_premats returned; the inverses of the
 * 128 x 128 diagonal blocks used by predict / simulate are rebuilt on the device (predictions agree with the
 * pre-save model to rounding, ~1e-13 relative, not bit for bit).  Reference: the slots of R/1_fgpm_Class.R:65-84 and
 * data/precalculated_Xfgpm_objects.rda, whose five models are restored this way in tests/. */
int fgp_problem_restore(fgp_problem* prob, int ker, double nugget, double sig2, const double* theta, const double* L,
                        const double* LInvY);

/* update() fast paths (the reference refits with fgpm() in every branch of R/6_updating.R).
 * fgp_update_var:  variance substitution (upd_subHypers, R/6_updating.R:510-524): L <- sqrt(s'/s) L, LInvY <- LInvY /
 *                  sqrt(s'/s) on the stored factor; outputs as fgp_premats (may be NULL).
 * fgp_update_y:    output substitution (upd_subData with sOut.sb only, R/6_updating.R:154-160): the factor is kept, only
 *                  L^-1 y is solved again, O(n^2).
 * fgp_premats_reuse: data addition / deletion / substitution with retained hyper-parameters (upd_add R/6_updating.R:325-500,
 *                  upd_del :7-89, upd_subData :91-323).  prob_new holds the updated data set (fgp_problem_create), prob_old the
 *                  fitted model.  The leading points both share (identical coordinates, same order), rounded down to a
 *                  multiple of 128, keep their block of the factor; the rest is assembled and eliminated against it.
 *                  n_reused (may be NULL) returns that count; 0 means a plain fgp_premats was done.  Result: the model of
 *                  prob_new with prob_old's ker, nugget, sig2, theta -- what fgpm() with known hypers would give. */
int fgp_update_var(fgp_problem* prob, double sig2_new, double* L_out, double* LInvY_out);
int fgp_update_y(fgp_problem* prob, const double* y_new, double* LInvY_out);
int fgp_premats_reuse(fgp_problem* prob_new, const fgp_problem* prob_old, double* L_out, double* LInvY_out, int* n_reused);

/* makePreds_*: (R/4_prediction_SF.R:2-21).  sNew m x ds, coefsNew[i] m x p[i] (already projected).
 * mean[m], sd[m]; detail = full additionally fills Ktp (n x m) and Kpp (m x m) when non-NULL. */
int fgp_predict(fgp_problem* prob, int m, const double* sNew, const double* const* coefsNew, int detail,
                double* mean, double* sd, double* Ktp, double* Kpp);

/* makeSims_*: (R/5_simulation_SF.R:3-23).  Z is matrix(rnorm(m*nsim), m, nsim) drawn by the caller with R's RNG
 * after set.seed(seed) (:13-14) so the R stream is untouched.  sims is nsim x m; mean/sd (m) may be NULL. */
int fgp_simulate(fgp_problem* prob, int m, const double* sNew, const double* const* coefsNew, int nsim,
                 const double* Z, double nugget_sm, double* sims, double* mean, double* sd);

/* getOut_loocv + Q2 (R/8_outilsStats.R:1-7, R/1_Xfgpm_Class.R:522-543), including quirk Q2 (nugget counted twice).
 * y_pre[n] (may be NULL), q2 scalar.  The result is kept with the model: asking again (getFitness, then the calibration
 * plot, R/7_plottingFunctions.R:9-11, 569-572, which recompute solve(R) for the same model) costs nothing until the
 * model changes. */
int fgp_loocv(fgp_problem* prob, double* y_pre, double* q2);

/* ---- batched candidate scoring for the ant-colony search ------------------------------------------ */
/* fitNtests_ACO / eval_loocv_ACO (R/3_ant_admin.R:592-648, 810-831): each candidate = one structure in the wire
 * format of formatSol_ACO (R/3_ant_admin.R:301-373): ints of length ds + 4*df + 1 --
 *   State X_i (1 on, 2 off) | per functional input: State, Dist (1 bygroup, 2 byindex), Dim, Basis (1 B-splines,
 *   2 PCA), -1 when off | Kernel (1 gauss, 2 matern5_2, 3 matern3_2).
 * ants is n_cand x (ds + 4*df + 1), one candidate per row (row-major).  fIn[j] is the raw n x k[j] curve matrix.
 * For every candidate: projection -> presample (u01 = the runif(d * max(n_presample, n_starts)) draws made by R in ant
 * order, so the R RNG stream is unchanged; column-major d x n_presample per candidate, concatenated; u01_off[c] = offset
 * of candidate c in doubles; d = its number of length-scales) ->
 * L-BFGS-B fit from the best n_starts points (lmm 5, factr 1e7, pgtol 0, maxit 100, central finite differences
 * with ndeps 1e-3 clipped at the box, as stats::optim does) -> preMats -> Q2loocv.
 * Outputs per candidate: fitness (Q2, NaN on crash), nll, sig2, theta (dmax per candidate, NaN padded), info (0, the
 * failing minor's order, or < 0: -1 bad structure, -4 more than dmax length-scales, -5 every start crashed).
 * Scheduling (option "fit_mode"): 1 (default) = lock-step -- ONE host thread per GPU of the context advances all
 * in-flight candidates together; every round the pending likelihood points of all of them (presample points,
 * finite-difference points of an L-BFGS-B iterate, the value at the optimum) form one heterogeneous batched launch
 * sequence, and the leave-one-out factorisations of the candidates that finished are batched the same way; options
 * "fit_points" (points per round and lane, default 1024 or what memory allows) and "fit_lanes" (rounds in flight per GPU).
 * 0 = one host thread per candidate, `fit_workers` (option, default 8) of them per GPU, each with its own streams and
 * workspaces.  Per-item results do not depend on the mode, the options or the number of GPUs; GPUs exchange nothing.
 * A failure of the library itself (CUDA error, out of memory) is returned as the call's error code, not reported as
 * crashed candidates. */
int fgp_fit_batch(fgp_ctx* ctx, int n, int ds, const double* sIn, int df, const int* k, const double* const* fIn,
                  const double* y, int n_cand, const int* ants, double nugget, int n_starts, int n_presample,
                  const double* u01, const long long* u01_off, int dmax, double* fitness, double* nll, double* sig2,
                  double* theta, int* info);

/* eval_houtv_ACO (R/3_ant_admin.R:662-802): the same with hold-out validation.  ind_vl is n_vl x n_rep (column-major,
 * 1-based row indices as in R's ind.vl matrix); every (candidate, replicate) pair is one work item: the model is fitted
 * on the rows not in column j (splitData, R/1_Xfgpm_Class.R:559-584; the projection basis is built from the training
 * rows) and scored with Q2hout = 1 - mean((y.vl - predict(model)$mean)^2) / mean((y.vl - mean(y.vl))^2)
 * (R/1_Xfgpm_Class.R:545-549).  Outputs and u01 offsets are per work item, candidate-major: item = c * n_rep + j.
 * Averaging over replicates and crash handling stay with the caller (R/3_ant_admin.R:712-728). */
int fgp_fit_batch_holdout(fgp_ctx* ctx, int n, int ds, const double* sIn, int df, const int* k, const double* const* fIn,
                          const double* y, int n_cand, const int* ants, double nugget, int n_starts, int n_presample,
                          const double* u01, const long long* u01_off, int dmax, int n_rep, int n_vl, const int* ind_vl,
                          double* fitness, double* nll, double* sig2, double* theta, int* info);

/* ---- instrumentation ------------------------------------------------------------------------------- */
/* timing of the last call on device 0 (CUDA events on the library's own streams), milliseconds / counts:
 *  [0] assembly kernel ms  [1] factorisation (all launches) ms  [2] trailing-update (DMMA) kernel ms summed
 *  [3] trailing-update launches  [4] trailing-update flops  [5] diagonal-block (POTF2 + inverse) kernel ms summed
 *  [6] panel-solve (TRSM as DMMA GEMM) ms summed  [7] kernel launches timed.  Per-launch event timing is only taken when fgp_set_option("profile", 1). */
int fgp_get_timing(fgp_ctx* ctx, double* out8);
/* device time (CUDA events on the launching stream) of the dataflow-Cholesky launch of the last fgp_nll_batch call on
 * device 0 -- ONE kernel (csrc/chol_dag.cu) factors all matrices of the call's last chunk; 0 when that call ran on the
 * launch-per-step engine (option "dag" 0, profile mode) */
int fgp_last_factor_ms(fgp_ctx* ctx, double* ms);
/* kernels of this library launched by the process so far */
long long fgp_launch_count(void);
/* work items of the process so far: which = 0 likelihood factorisations (n^3/3 flops each), 1 leave-one-out
 * factorisations (Cholesky + L^-T, 2 n^3/3), 2 preMats factorisations, 3 candidates scored by fgp_fit_batch */
long long fgp_work_count(int which);
/* device-side stopwatch over all of the library's streams on device 0 (CUDA events): idx, i, j in 0..3 */
int fgp_mark(fgp_ctx* ctx, int idx);
int fgp_elapsed_ms(fgp_ctx* ctx, int i, int j, double* ms);
/* options: "profile" (0/1), "lookahead" (0/1), "streams" (1..8 concurrent factorisations per GPU),
 * "outer" (column tiles of 128 per outer panel; the trailing update runs with K = outer*128), "outer_single" (the same for the look-ahead
 * path of a single factorisation; 0 = chosen by size: 2 below n = 12288, 8 below 24576, else 16), "fit_workers" (1..16),
 * "fit_blocking" (host waits of the fgp_fit_batch worker contexts: 0 spin, 1 sleep on a blocking event, 2 poll and yield;
 * default -1 = yield when workers x visible GPUs exceed 3/4 of the host cores, else spin; the lock-step driver spins unless > 0),
 * "fit_mode", "fit_points", "fit_lanes" (see fgp_fit_batch), "predict_reserve" (rows kept free under a stored factor),
 * "blocking_sync" (0/1/2, default 0: the same for this context's own calls),
 * "graphs" (0/1, default 0: batched likelihood calls at n < 2048 are captured as a CUDA graph the second time a shape is
 * seen on a problem and replayed afterwards -- one driver call per evaluation batch instead of ~30),
 * "dag" (0/1, default 1: the persistent dataflow kernel csrc/chol_dag.cu -- one launch factors all matrices of a
 * likelihood batch / preMats / a round of the candidate driver (any n), its solve mode eliminates the new rows of
 * predict / simulate / fgp_update_y (from 16 tiles on) and the identity rows of fgp_loocv (from n = 4096 on); results are
 * bit-identical to the launch-per-step engine it replaces, which 0 selects everywhere),
 * "fused_assembly", "pdl", "gemm_tma" (0/1, default 1: the batch-fused assembly kernel, programmatic dependent launch,
 * TMA operand path of the DMMA GEMM), "gemm_flat" (default 8: batched GEMM launches with at least this many tiles per SM
 * and K <= 1024 run one persistent CTA per SM over all (tile, matrix) pairs; 0 = never), "la_r8", "la_r4" (look-ahead
 * panel width thresholds in remaining tile columns), "host_threads" (host threads of the chunked L / K.tp transfers,
 * default 0 = min(16, cores)), "host_thp" (0/1, default 1: MADV_HUGEPAGE hint on host outputs of 64 MB and more),
 * "out_prezeroed" (0/1, default 0: the caller guarantees that every L_out it passes is zero-filled on entry, e.g. from
 * calloc; the library then writes the lower triangle only instead of also clearing the upper one). */
int fgp_set_option(fgp_ctx* ctx, const char* name, double value);
/* measured FP64 peaks on device 0 for the roofline denominators: register-resident DMMA (m8n8k4) loop and
 * DFMA loop, TFLOP/s; and a device copy, GB/s. */
int fgp_measure_peaks(fgp_ctx* ctx, double* dmma_tflops, double* dfma_tflops, double* copy_gbs);

#ifdef __cplusplus
}
#endif
#endif /* FUNGP_B200_H */
