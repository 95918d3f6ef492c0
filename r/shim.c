/* r/shim.c -- thin .Call shim between R and libgpb200.so.
 *
 * NOT COMPILED IN THE BUILD CONTAINER (no R, no Rinternals.h there); build where R exists with
 *     R CMD SHLIB r/shim.c -I include -L gp_b200/lib -lgpb200 -o r/gpb200_r.so
 * (see r/build.sh).  Every entry point only converts SEXPs to plain pointers, calls the C ABI of
 * include/gpb200.h and turns a non-zero status into an R error -- the arithmetic is in the CUDA
 * library.  It replaces the wrapper Rcpp attributes generate for covariance.cpp:8-9
 * (`extern "C" SEXP sourceCpp_N_rbf_cov_chol(SEXP x1SEXP, SEXP l_SEXP)`) and gives the R kernel and
 * conditioning functions of R/kernels.R, derivative_kernels.R, R/ode_gp_library.R one call per
 * matrix instead of one closure call per element.
 *
 * R is single-threaded: all entry points run on the main R thread and synchronise before returning
 * (host-pointer mode of the ABI).  Outputs are R-allocated and PROTECTed while being filled.
 */
#include <stdlib.h>
#include <string.h>

#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "gpb200.h"

static gpb200_handle_t g_h = NULL;

static gpb200_handle_t handle(void) {
  if (!g_h) {
    int dev = 0;
    const char *e = getenv("GPB200_DEVICE");
    if (e) dev = atoi(e);
    int rc = gpb200_create(&g_h, dev);
    if (rc != 0) Rf_error("gpb200: no usable B200 GPU (gpb200_create returned %d); there is no CPU fallback", rc);
  }
  return g_h;
}

static void check(int rc, const char *where) {
  if (rc < 0) Rf_error("%s failed (%d): %s", where, rc, gpb200_last_error(g_h));
  if (rc > 0) Rf_error("%s: matrix is not positive definite (first non-positive pivot at %d)", where, rc);
}

/* rbf_cov_chol(x1, l_) -> list(L =, dLdl =)          [covariance.cpp:8-47] */
SEXP gp_rbf_cov_chol(SEXP x1, SEXP l_) {
  const int n = LENGTH(x1);
  SEXP L = PROTECT(Rf_allocMatrix(REALSXP, n, n));
  SEXP dL = PROTECT(Rf_allocMatrix(REALSXP, n, n));
  check(gpb200_rbf_cov_chol(handle(), n, REAL(x1), Rf_asReal(l_), REAL(L), REAL(dL)), "rbf_cov_chol");
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_VECTOR_ELT(out, 0, L);
  SET_VECTOR_ELT(out, 1, dL);
  SET_STRING_ELT(nm, 0, Rf_mkChar("L"));
  SET_STRING_ELT(nm, 1, Rf_mkChar("dLdl"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(4);
  return out;
}

/* approx_L(l, lp, Ls, dLdls)                         [covariance.cpp:49-96] */
SEXP gp_approx_L(SEXP l, SEXP lp, SEXP Ls, SEXP dLdls) {
  const int P = LENGTH(lp);
  if (LENGTH(Ls) != P || LENGTH(dLdls) != P) Rf_error("approx_L: lp, Ls, dLdls must have equal length");
  const int n = Rf_nrows(VECTOR_ELT(Ls, 0));
  const double **a = (const double **)R_alloc(P, sizeof(double *));
  const double **b = (const double **)R_alloc(P, sizeof(double *));
  for (int i = 0; i < P; i++) { a[i] = REAL(VECTOR_ELT(Ls, i)); b[i] = REAL(VECTOR_ELT(dLdls, i)); }
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, n));
  check(gpb200_approx_L(handle(), n, Rf_asReal(l), P, REAL(lp), a, b, REAL(out)), "approx_L");
  UNPROTECT(1);
  return out;
}

/* all P tables of a length-scale grid in one batched call -> list(Ls = list(...), dLdls = list(...))
 * (data block of models/cubic_interpolated_gp.stan:11-12; interpolated_gp.stan:15-21) */
SEXP gp_rbf_cov_chol_grid(SEXP x1, SEXP lp) {
  const int n = LENGTH(x1), P = LENGTH(lp);
  double *L = (double *)R_alloc((size_t)P * n * n, sizeof(double));
  double *dL = (double *)R_alloc((size_t)P * n * n, sizeof(double));
  int *info = (int *)R_alloc(P, sizeof(int));
  check(gpb200_rbf_cov_chol_batched(handle(), n, REAL(x1), P, REAL(lp), L, dL, info), "rbf_cov_chol_grid");
  for (int q = 0; q < P; q++)
    if (info[q] > 0) Rf_error("rbf_cov_chol_grid: table %d is not positive definite (pivot %d)", q + 1, info[q]);
  SEXP Ls = PROTECT(Rf_allocVector(VECSXP, P));
  SEXP dLs = PROTECT(Rf_allocVector(VECSXP, P));
  for (int q = 0; q < P; q++) {
    SEXP a = PROTECT(Rf_allocMatrix(REALSXP, n, n));
    SEXP b = PROTECT(Rf_allocMatrix(REALSXP, n, n));
    memcpy(REAL(a), L + (size_t)q * n * n, sizeof(double) * (size_t)n * n);
    memcpy(REAL(b), dL + (size_t)q * n * n, sizeof(double) * (size_t)n * n);
    SET_VECTOR_ELT(Ls, q, a);
    SET_VECTOR_ELT(dLs, q, b);
    UNPROTECT(2);
  }
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_VECTOR_ELT(out, 0, Ls); SET_VECTOR_ELT(out, 1, dLs);
  SET_STRING_ELT(nm, 0, Rf_mkChar("Ls")); SET_STRING_ELT(nm, 1, Rf_mkChar("dLdls"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(4);
  return out;
}

/* approx_L(M, scale, xt, sigma, l) of models/westbrook.stan:2-30 / bH of spectral_test.R:6-27 */
SEXP gp_approx_L_basis(SEXP M, SEXP scale, SEXP x, SEXP sigma, SEXP l) {
  const int n = LENGTH(x), m = Rf_asInteger(M);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, m));
  check(gpb200_approx_L_basis(handle(), n, m, Rf_asReal(scale), REAL(x), Rf_asReal(sigma), Rf_asReal(l), REAL(out),
                              n > 0 ? n : 1), "approx_L_basis");
  UNPROTECT(1);
  return out;
}

/* L = chol(cov_exp_quad(x, alpha, rho) + diag_add I) and dL/d(alpha | rho): the latent models' Cholesky
 * with its tangent (exact_gp.stan:17-25, fit_full_gp.stan:18-26) */
SEXP gp_se_chol_tangent(SEXP x, SEXP alpha, SEXP rho, SEXP diag_add, SEXP wrt) {
  const int n = LENGTH(x);
  SEXP L = PROTECT(Rf_allocMatrix(REALSXP, n, n));
  SEXP dL = PROTECT(Rf_allocMatrix(REALSXP, n, n));
  check(gpb200_se_chol_tangent(handle(), n, REAL(x), Rf_asReal(alpha), Rf_asReal(rho), Rf_asReal(diag_add),
                               Rf_asInteger(wrt), REAL(L), REAL(dL)), "se_chol_tangent");
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_VECTOR_ELT(out, 0, L); SET_VECTOR_ELT(out, 1, dL);
  SET_STRING_ELT(nm, 0, Rf_mkChar("L")); SET_STRING_ELT(nm, 1, Rf_mkChar("dL"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(4);
  return out;
}

/* outer(x, y, kernel) in one call: kind per include/gpb200.h  [derivative_kernels.R, R/kernels.R] */
SEXP gp_gram_outer(SEXP kind, SEXP x, SEXP y, SEXP amp2, SEXP l) {
  const int n = LENGTH(x), m = LENGTH(y);
  SEXP K = PROTECT(Rf_allocMatrix(REALSXP, n, m));
  check(gpb200_gram_outer(handle(), Rf_asInteger(kind), n, m, REAL(x), REAL(y), Rf_asReal(amp2), Rf_asReal(l),
                          REAL(K), n > 0 ? n : 1), "gram_outer");
  UNPROTECT(1);
  return K;
}

/* element-wise kernel (vectorised R closure semantics) */
SEXP gp_kernel_eval(SEXP kind, SEXP tj, SEXP tk, SEXP amp2, SEXP l) {
  const R_xlen_t len = XLENGTH(tj);
  if (XLENGTH(tk) != len) Rf_error("kernel_eval: tj and tk must have the same length (recycle in R first)");
  SEXP out = PROTECT(Rf_allocVector(REALSXP, len));
  check(gpb200_kernel_eval(handle(), Rf_asInteger(kind), (long long)len, REAL(tj), REAL(tk), Rf_asReal(amp2),
                           Rf_asReal(l), REAL(out)), "kernel_eval");
  UNPROTECT(1);
  return out;
}

/* QQard(X, Y, phi)                                   [R/kernels.R:19] */
SEXP gp_gram_ard(SEXP X, SEXP Y, SEXP alpha, SEXP rho) {
  const int n = Rf_nrows(X), D = Rf_ncols(X), m = Rf_nrows(Y);
  if (Rf_ncols(Y) != D) Rf_error("QQard: X and Y must have the same number of columns");
  double *r = (double *)R_alloc(D, sizeof(double));
  for (int d = 0; d < D; d++) r[d] = REAL(rho)[LENGTH(rho) == 1 ? 0 : d];
  SEXP K = PROTECT(Rf_allocMatrix(REALSXP, n, m));
  check(gpb200_gram_ard(handle(), n, m, D, REAL(X), n, REAL(Y), m, Rf_asReal(alpha), r, REAL(K), n), "gram_ard");
  UNPROTECT(1);
  return K;
}

/* joint derivative covariance                        [R/ode_gp_library.R:29-30; design_notes.Rmd] */
SEXP gp_gram_deriv(SEXP t, SEXP alpha, SEXP rho, SEXP nblocks, SEXP noise, SEXP jitter, SEXP quirk) {
  const int n = LENGTH(t), nb = Rf_asInteger(nblocks), N = n * nb;
  SEXP K = PROTECT(Rf_allocMatrix(REALSXP, N, N));
  check(gpb200_gram_deriv(handle(), n, REAL(t), Rf_asReal(alpha), Rf_asReal(rho), nb, REAL(noise),
                          Rf_asReal(jitter), Rf_asInteger(quirk), REAL(K), N > 0 ? N : 1), "gram_deriv");
  UNPROTECT(1);
  return K;
}

/* chol(K) lower                                      [spectral_test.R:32; cholesky_decompose] */
SEXP gp_potrf(SEXP K) {
  const int n = Rf_nrows(K);
  SEXP L = PROTECT(Rf_duplicate(K));
  check(gpb200_potrf(handle(), n, REAL(L), n > 0 ? n : 1), "cholesky_decompose");
  UNPROTECT(1);
  return L;
}

/* LML + gradient for B draws: theta is a 3 x B matrix (alpha, rho, sigma per column) */
SEXP gp_lml_grad_draws(SEXP x, SEXP y, SEXP theta, SEXP jitter) {
  const int n = LENGTH(x), B = Rf_ncols(theta);
  SEXP lml = PROTECT(Rf_allocVector(REALSXP, B));
  SEXP grad = PROTECT(Rf_allocMatrix(REALSXP, 3, B));
  SEXP info = PROTECT(Rf_allocVector(INTSXP, B));
  check(gpb200_lml_grad_batched(handle(), n, B, REAL(x), 0, REAL(y), 0, REAL(theta), Rf_asReal(jitter), 1,
                                REAL(lml), REAL(grad), INTEGER(info)), "lml_grad_draws");
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 3));
  SET_VECTOR_ELT(out, 0, lml); SET_VECTOR_ELT(out, 1, grad); SET_VECTOR_ELT(out, 2, info);
  SET_STRING_ELT(nm, 0, Rf_mkChar("lml")); SET_STRING_ELT(nm, 1, Rf_mkChar("grad")); SET_STRING_ELT(nm, 2, Rf_mkChar("info"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(5);
  return out;
}

/* LML + gradient of a GP observed through derivative orders order0 .. order0+nblocks-1 on the grid t
 * (gpderivs.py:62-83 is order0 = 1, nblocks = 1; design_notes.Rmd:25-46 is 0, 3); theta is a
 * (2 + nblocks) x B matrix (alpha, rho, noise[nblocks] per column), y the stacked observations */
SEXP gp_lml_grad_deriv_draws(SEXP t, SEXP y, SEXP theta, SEXP order0, SEXP jitter) {
  const int n = LENGTH(t), B = Rf_ncols(theta), nb = Rf_nrows(theta) - 2;
  if (nb < 1 || LENGTH(y) != n * nb) Rf_error("gp_lml_grad_deriv_draws: y must hold n * (nrow(theta) - 2) values");
  SEXP lml = PROTECT(Rf_allocVector(REALSXP, B));
  SEXP grad = PROTECT(Rf_allocMatrix(REALSXP, 2 + nb, B));
  SEXP info = PROTECT(Rf_allocVector(INTSXP, B));
  check(gpb200_lml_grad_deriv_batched(handle(), n, Rf_asInteger(order0), nb, B, REAL(t), 0, REAL(y), 0, REAL(theta),
                                      Rf_asReal(jitter), 1, REAL(lml), REAL(grad), INTEGER(info)), "lml_grad_deriv_draws");
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 3));
  SET_VECTOR_ELT(out, 0, lml); SET_VECTOR_ELT(out, 1, grad); SET_VECTOR_ELT(out, 2, info);
  SET_STRING_ELT(nm, 0, Rf_mkChar("lml")); SET_STRING_ELT(nm, 1, Rf_mkChar("grad")); SET_STRING_ELT(nm, 2, Rf_mkChar("info"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(5);
  return out;
}

/* MASS::mvrnorm(n, mu, Sigma) with the device generator (pendulum_fit.R:253): n x length(mu) matrix */
SEXP gp_mvrnorm(SEXP n_, SEXP mu, SEXP Sigma, SEXP seed) {
  const int nd = Rf_asInteger(n_), m = Rf_nrows(Sigma);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, nd, m));
  check(gpb200_mvrnorm(handle(), nd, m, REAL(mu), REAL(Sigma), m, 0.0, (unsigned long long)Rf_asReal(seed), REAL(out), nd > 0 ? nd : 1),
        "mvrnorm");
  UNPROTECT(1);
  return out;
}

/* mu = Ks (K + s2 I)^-1 y ; cov = Kss - Ks (K + s2 I)^-1 Ks^T + jitter I   [pendulum_fit.R:242-251] */
SEXP gp_condition(SEXP K, SEXP Ks, SEXP Kss, SEXP y, SEXP noise_var, SEXP jitter) {
  const int n = Rf_nrows(K), m = Rf_nrows(Ks);
  SEXP mu = PROTECT(Rf_allocVector(REALSXP, m));
  SEXP cov = PROTECT(Rf_allocMatrix(REALSXP, m, m));
  check(gpb200_gp_condition(handle(), n, m, REAL(K), n, REAL(Ks), m, REAL(Kss), m, REAL(y), Rf_asReal(noise_var),
                            Rf_asReal(jitter), REAL(mu), REAL(cov), m), "gp_condition");
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_VECTOR_ELT(out, 0, mu); SET_VECTOR_ELT(out, 1, cov);
  SET_STRING_ELT(nm, 0, Rf_mkChar("mu")); SET_STRING_ELT(nm, 1, Rf_mkChar("cov"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(4);
  return out;
}

/* condMVN for the reference's block layout (given block first)   [R/ode_gp_library.R:17,32] */
SEXP gp_cond_mvn(SEXP mean, SEXP sigma, SEXP ng_, SEXP x_given) {
  const int N = Rf_nrows(sigma), ng = Rf_asInteger(ng_), nd = N - ng;
  SEXP cm = PROTECT(Rf_allocVector(REALSXP, nd));
  SEXP cv = PROTECT(Rf_allocMatrix(REALSXP, nd, nd));
  check(gpb200_cond_mvn(handle(), ng, nd, REAL(mean), REAL(sigma), N, REAL(x_given), REAL(cm), REAL(cv), nd), "condMVN");
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_VECTOR_ELT(out, 0, cm); SET_VECTOR_ELT(out, 1, cv);
  SET_STRING_ELT(nm, 0, Rf_mkChar("condMean")); SET_STRING_ELT(nm, 1, Rf_mkChar("condVar"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(4);
  return out;
}

static const R_CallMethodDef call_methods[] = {
    {"gp_rbf_cov_chol", (DL_FUNC)&gp_rbf_cov_chol, 2}, {"gp_approx_L", (DL_FUNC)&gp_approx_L, 4},
    {"gp_gram_outer", (DL_FUNC)&gp_gram_outer, 5},     {"gp_kernel_eval", (DL_FUNC)&gp_kernel_eval, 5},
    {"gp_gram_ard", (DL_FUNC)&gp_gram_ard, 4},         {"gp_gram_deriv", (DL_FUNC)&gp_gram_deriv, 7},
    {"gp_potrf", (DL_FUNC)&gp_potrf, 1},               {"gp_lml_grad_draws", (DL_FUNC)&gp_lml_grad_draws, 4},
    {"gp_condition", (DL_FUNC)&gp_condition, 6},       {"gp_cond_mvn", (DL_FUNC)&gp_cond_mvn, 4},
    {"gp_rbf_cov_chol_grid", (DL_FUNC)&gp_rbf_cov_chol_grid, 2}, {"gp_approx_L_basis", (DL_FUNC)&gp_approx_L_basis, 5},
    {"gp_se_chol_tangent", (DL_FUNC)&gp_se_chol_tangent, 5},
    {"gp_lml_grad_deriv_draws", (DL_FUNC)&gp_lml_grad_deriv_draws, 5}, {"gp_mvrnorm", (DL_FUNC)&gp_mvrnorm, 4},
    {NULL, NULL, 0}};

void R_init_gpb200_r(DllInfo *dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
