/* registration of the .Call entry points: NAMESPACE gains  useDynLib(funGp, .registration = TRUE)  */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

extern SEXP fgpR_ctx_create(SEXP);
extern SEXP fgpR_project(SEXP, SEXP, SEXP, SEXP, SEXP);
extern SEXP fgpR_problem_create(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
extern SEXP fgpR_dist_max(SEXP);
extern SEXP fgpR_dist_matrix(SEXP, SEXP);
extern SEXP fgpR_assemble(SEXP, SEXP, SEXP, SEXP);
extern SEXP fgpR_nll_batch(SEXP, SEXP, SEXP, SEXP, SEXP);
extern SEXP fgpR_premats(SEXP, SEXP, SEXP, SEXP, SEXP);
extern SEXP fgpR_predict(SEXP, SEXP, SEXP, SEXP, SEXP);
extern SEXP fgpR_simulate(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
extern SEXP fgpR_loocv(SEXP);
extern SEXP fgpR_set_y(SEXP, SEXP);
extern SEXP fgpR_fit_batch(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
extern SEXP fgpR_ptr_valid(SEXP);
extern SEXP fgpR_restore(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
extern SEXP fgpR_update_var(SEXP, SEXP);
extern SEXP fgpR_update_y(SEXP, SEXP);
extern SEXP fgpR_premats_reuse(SEXP, SEXP);

static const R_CallMethodDef CallEntries[] = {
    {"fgpR_ctx_create", (DL_FUNC)&fgpR_ctx_create, 1},       {"fgpR_project", (DL_FUNC)&fgpR_project, 5},
    {"fgpR_problem_create", (DL_FUNC)&fgpR_problem_create, 6}, {"fgpR_dist_max", (DL_FUNC)&fgpR_dist_max, 1},
    {"fgpR_dist_matrix", (DL_FUNC)&fgpR_dist_matrix, 2},     {"fgpR_assemble", (DL_FUNC)&fgpR_assemble, 4},
    {"fgpR_nll_batch", (DL_FUNC)&fgpR_nll_batch, 5},         {"fgpR_premats", (DL_FUNC)&fgpR_premats, 5},
    {"fgpR_predict", (DL_FUNC)&fgpR_predict, 5},             {"fgpR_simulate", (DL_FUNC)&fgpR_simulate, 6},
    {"fgpR_loocv", (DL_FUNC)&fgpR_loocv, 1},                 {"fgpR_set_y", (DL_FUNC)&fgpR_set_y, 2},
    {"fgpR_fit_batch", (DL_FUNC)&fgpR_fit_batch, 12},
    {"fgpR_ptr_valid", (DL_FUNC)&fgpR_ptr_valid, 1},         {"fgpR_restore", (DL_FUNC)&fgpR_restore, 7},
    {"fgpR_update_var", (DL_FUNC)&fgpR_update_var, 2},       {"fgpR_update_y", (DL_FUNC)&fgpR_update_y, 2},
    {"fgpR_premats_reuse", (DL_FUNC)&fgpR_premats_reuse, 2},
    {NULL, NULL, 0}};

void R_init_funGp(DllInfo* dll) {
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
