/* Plain-C CPU restatement of the GP hot path of bbbales2/gp -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  "Parity unpinned" for the Stan-Math rows: the reference holds no golden
 * output for them and Stan Math / Eigen / Rcpp are not vendored (see oracle/gp_oracle.py header).
 * This file is the "second opinion" next to the NumPy oracle: same semantics, independent code,
 * Stan-Math-like loop order (lower-triangle fill + mirror, column LLT, forward substitution).
 *
 * All matrices are column-major (R / Eigen default), double precision, no fast-math.
 * Citations are relative to /root/reference.
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=native -fPIC -shared -pthread)
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define IDX(i, j, ld) ((size_t)(i) + (size_t)(j) * (size_t)(ld))

/* a4: cov_exp_quad (models/fit_hyperparameters.stan:19): diagonal = alpha^2 exactly, lower
 * triangle alpha^2 exp(-0.5 d^2 / rho^2), mirrored. */
void oracle_cov_exp_quad(int n, const double *x, double alpha, double rho, double *K) {
  const double a2 = alpha * alpha, nhr = -0.5 / (rho * rho);
  for (int j = 0; j < n; j++) {
    K[IDX(j, j, n)] = a2;
    for (int i = j + 1; i < n; i++) {
      const double d = x[i] - x[j];
      const double v = a2 * exp(d * d * nhr);
      K[IDX(i, j, n)] = v;
      K[IDX(j, i, n)] = v;
    }
  }
}

/* a5: diagonal add (fit_hyperparameters.stan:21-24 and the jitter sites of SURVEY 8a row a5). */
void oracle_add_diag(int n, double *K, double c) {
  for (int i = 0; i < n; i++) K[IDX(i, i, n)] += c;
}

/* a6: cholesky_decompose (fit_hyperparameters.stan:25; covariance.cpp:29).  In-place lower LLT,
 * blocked right-looking (block 64) so that the host baseline is not artificially slow; returns
 * LAPACK-style info (k>0: first non-positive pivot, 1-based).  The strict upper triangle is
 * zeroed like Stan/Eigen's matrixL(). */
int oracle_llt(int n, double *A) {
  const int NB = 64;
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int kb = (n - k0 < NB) ? n - k0 : NB;
    /* factor the (n-k0) x kb panel: diagonal block by the unblocked routine applied to the tall
     * panel (it also scales the rows below) */
    {
      double *P = &A[IDX(k0, k0, n)];
      const int m = n - k0;
      for (int j = 0; j < kb; j++) {
        double s = P[IDX(j, j, n)];
        if (!(s > 0.0)) return k0 + j + 1;
        const double d = sqrt(s);
        P[IDX(j, j, n)] = d;
        const double inv = 1.0 / d;
        for (int i = j + 1; i < m; i++) P[IDX(i, j, n)] *= inv;
        for (int c = j + 1; c < kb; c++) {
          const double lcj = P[IDX(c, j, n)];
          double *restrict cc = &P[IDX(0, c, n)];
          const double *restrict cj = &P[IDX(0, j, n)];
          for (int i = c; i < m; i++) cc[i] -= cj[i] * lcj;
        }
      }
    }
    /* trailing update A22 -= L21 L21^T (lower part), 4 rank-1 terms per pass over a column */
    const int r0 = k0 + kb;
    for (int j = r0; j < n; j++) {
      double *restrict cj = &A[IDX(0, j, n)];
      int k = k0;
      for (; k + 3 < k0 + kb; k += 4) {
        const double b0 = A[IDX(j, k, n)], b1 = A[IDX(j, k + 1, n)], b2 = A[IDX(j, k + 2, n)], b3 = A[IDX(j, k + 3, n)];
        const double *restrict a0 = &A[IDX(0, k, n)], *restrict a1 = &A[IDX(0, k + 1, n)];
        const double *restrict a2 = &A[IDX(0, k + 2, n)], *restrict a3 = &A[IDX(0, k + 3, n)];
        for (int i = j; i < n; i++) cj[i] -= a0[i] * b0 + a1[i] * b1 + a2[i] * b2 + a3[i] * b3;
      }
      for (; k < k0 + kb; k++) {
        const double b0 = A[IDX(j, k, n)];
        const double *restrict a0 = &A[IDX(0, k, n)];
        for (int i = j; i < n; i++) cj[i] -= a0[i] * b0;
      }
    }
  }
  for (int j = 1; j < n; j++)
    for (int i = 0; i < j; i++) A[IDX(i, j, n)] = 0.0;
  return 0;
}

/* mdivide_left_tri_low: solve L z = b in place (inside multi_normal_cholesky,
 * fit_hyperparameters.stan:31).  Column-oriented forward substitution. */
void oracle_trsv_lower(int n, const double *L, double *b) {
  for (int j = 0; j < n; j++) {
    const double v = b[j] / L[IDX(j, j, n)];
    b[j] = v;
    const double *restrict c = &L[IDX(0, j, n)];
    for (int i = j + 1; i < n; i++) b[i] -= c[i] * v;
  }
}

/* solve L^T a = z in place (dot-product form, contiguous column reads). */
void oracle_trsv_lower_t(int n, const double *L, double *b) {
  for (int j = n - 1; j >= 0; j--) {
    const double *restrict c = &L[IDX(0, j, n)];
    double s = b[j];
    for (int i = j + 1; i < n; i++) s -= c[i] * b[i];
    b[j] = s / c[j];
  }
}

/* a8: f = L z, lower-triangular (exact_gp.stan:25). */
void oracle_trmv_lower(int n, const double *L, const double *z, double *f) {
  for (int i = 0; i < n; i++) f[i] = 0.0;
  for (int j = 0; j < n; j++) {
    const double zj = z[j];
    const double *restrict c = &L[IDX(0, j, n)];
    for (int i = j; i < n; i++) f[i] += c[i] * zj;
  }
}

/* a7: multi_normal_cholesky_lpdf with mu = 0 (fit_hyperparameters.stan:31). */
double oracle_mvn_chol_lpdf(int n, const double *y, const double *L, int drop_constants) {
  double *z = (double *)malloc(sizeof(double) * (size_t)n);
  memcpy(z, y, sizeof(double) * (size_t)n);
  oracle_trsv_lower(n, L, z);
  double q = 0.0, ld = 0.0;
  for (int i = 0; i < n; i++) { q += z[i] * z[i]; ld += log(L[IDX(i, i, n)]); }
  free(z);
  double lp = -ld - 0.5 * q;
  if (!drop_constants) lp -= 0.5 * n * log(2.0 * M_PI);
  return lp;
}

/* CS-A: LML and gradient w.r.t. theta = (alpha, rho, sigma) for K = cov_exp_quad + (sigma^2+jitter) I
 * (models/fit_hyperparameters.stan:18-31 + the reverse sweep; formulas SURVEY Appendix B).
 * work: caller-provided scratch of 2*n*n doubles (or NULL to malloc).  Returns info. */
int oracle_lml_grad(int n, const double *x, const double *y, const double *theta, double jitter,
                    double *lml, double *grad, double *work) {
  const double alpha = theta[0], rho = theta[1], sigma = theta[2];
  const size_t nn = (size_t)n * (size_t)n;
  double *own = NULL;
  if (!work) { own = (double *)malloc(sizeof(double) * 2 * nn); work = own; }
  double *L = work, *W = work + nn;
  oracle_cov_exp_quad(n, x, alpha, rho, L);
  oracle_add_diag(n, L, sigma * sigma + jitter);
  const int info = oracle_llt(n, L);
  if (info) { if (own) free(own); *lml = NAN; grad[0] = grad[1] = grad[2] = NAN; return info; }
  double *z = (double *)malloc(sizeof(double) * 2 * (size_t)n), *a = z + n;
  memcpy(z, y, sizeof(double) * (size_t)n);
  oracle_trsv_lower(n, L, z);
  double q = 0.0, ld = 0.0;
  for (int i = 0; i < n; i++) { q += z[i] * z[i]; ld += log(L[IDX(i, i, n)]); }
  *lml = -0.5 * n * log(2.0 * M_PI) - ld - 0.5 * q;
  memcpy(a, z, sizeof(double) * (size_t)n);
  oracle_trsv_lower_t(n, L, a);
  /* W = L^-1, column by column (forward substitution on unit vectors; column c is zero above c) */
  memset(W, 0, sizeof(double) * nn);
  for (int c = 0; c < n; c++) {
    double *restrict w = &W[IDX(0, c, n)];
    w[c] = 1.0;
    for (int j = c; j < n; j++) {
      const double v = w[j] / L[IDX(j, j, n)];
      w[j] = v;
      const double *restrict lc = &L[IDX(0, j, n)];
      for (int i = j + 1; i < n; i++) w[i] -= lc[i] * v;
    }
  }
  /* contraction 0.5 tr((a a^T - K^-1) dK/dtheta), K^-1_ij = sum_{k>=i} W_ki W_kj (i >= j) */
  const double a2 = alpha * alpha, nhr = -0.5 / (rho * rho), rho3 = rho * rho * rho;
  double s_se = 0.0, s_d2 = 0.0, tr = 0.0, aa = 0.0;
  for (int j = 0; j < n; j++) {
    for (int i = j; i < n; i++) {
      const double *restrict wi = &W[IDX(0, i, n)], *restrict wj = &W[IDX(0, j, n)];
      double g = 0.0;
      for (int k = i; k < n; k++) g += wi[k] * wj[k];
      const double m = a[i] * a[j] - g;
      if (i == j) { tr += g; aa += a[i] * a[i]; s_se += m; }
      else {
        const double d = x[i] - x[j];
        const double e = exp(d * d * nhr);
        s_se += 2.0 * m * e;
        s_d2 += 2.0 * m * e * d * d;
      }
    }
  }
  grad[0] = alpha * s_se;                 /* 0.5 * sum M * 2 alpha e        */
  grad[1] = 0.5 * a2 * s_d2 / rho3;       /* 0.5 * sum M * alpha^2 e d^2/rho^3 */
  grad[2] = sigma * (aa - tr);            /* 0.5 * tr(M) * 2 sigma          */
  free(z);
  if (own) free(own);
  return 0;
}

/* Independent draws over P host threads: mirrors mclapply(s_list, ..., mc.cores = P)
 * (pendulum_fit.R:268) / rstan cores (pendulum_fit.R:206).  out = B x 5 (lml, g[3], info). */
typedef struct { int n, B, tid, nthreads; const double *x, *y, *theta; double jitter; double *out; } job_t;

static void *worker(void *arg) {
  job_t *jb = (job_t *)arg;
  double *work = (double *)malloc(sizeof(double) * 2 * (size_t)jb->n * (size_t)jb->n);
  for (int b = jb->tid; b < jb->B; b += jb->nthreads) {
    double *o = jb->out + 5 * (size_t)b;
    const int info = oracle_lml_grad(jb->n, jb->x, jb->y, jb->theta + 3 * (size_t)b, jb->jitter, &o[0], &o[1], work);
    o[4] = (double)info;
  }
  free(work);
  return NULL;
}

void oracle_lml_grad_draws(int n, const double *x, const double *y, int B, const double *theta,
                           double jitter, int nthreads, double *out) {
  if (nthreads < 1) nthreads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (job_t){n, B, t, nthreads, x, y, theta, jitter, out};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
}

/* a1-a3: rbf_cov_chol (covariance.cpp:9-47), literal: full-square Gram in (value, tangent) pairs
 * seeded on l (:13,17-21), + 1e-10 (:23-25), LLT carried on the pairs (:29), unpack (:31-39). */
int oracle_rbf_cov_chol(int n, const double *x1, double l, double *L, double *dLdl) {
  const double l2 = 2.0 * l * l, l3 = l * l * l;
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) {
      const double d = x1[i] - x1[j];
      const double v = exp(-(d * d) / l2);
      L[IDX(i, j, n)] = v;
      dLdl[IDX(i, j, n)] = v * d * d / l3;
    }
  for (int i = 0; i < n; i++) L[IDX(i, i, n)] += 1e-10;
  /* column LLT on duals: value part in L, tangent part in dLdl */
  for (int j = 0; j < n; j++) {
    for (int k = 0; k < j; k++) {
      const double vjk = L[IDX(j, k, n)], tjk = dLdl[IDX(j, k, n)];
      for (int i = j; i < n; i++) {
        const double vik = L[IDX(i, k, n)], tik = dLdl[IDX(i, k, n)];
        L[IDX(i, j, n)] -= vik * vjk;
        dLdl[IDX(i, j, n)] -= vik * tjk + tik * vjk;
      }
    }
    const double sv = L[IDX(j, j, n)], st = dLdl[IDX(j, j, n)];
    if (!(sv > 0.0)) return j + 1;
    const double dv = sqrt(sv), dt = 0.5 * st / dv;
    L[IDX(j, j, n)] = dv;
    dLdl[IDX(j, j, n)] = dt;
    for (int i = j + 1; i < n; i++) {
      const double v = L[IDX(i, j, n)] / dv;
      dLdl[IDX(i, j, n)] = (dLdl[IDX(i, j, n)] - v * dt) / dv;
      L[IDX(i, j, n)] = v;
    }
  }
  for (int j = 1; j < n; j++)
    for (int i = 0; i < j; i++) { L[IDX(i, j, n)] = 0.0; dLdl[IDX(i, j, n)] = 0.0; }
  return 0;
}

/* a9: the nine derivative kernels, element-wise (derivative_kernels.R:39-73), times amp2.
 * kind: 0 QQ 1 QR 2 RQ 3 RR 4 QT 5 TQ 6 RT 7 TR 8 TT */
double oracle_deriv_kernel(int kind, double tj, double tk, double l) {
  if (kind == 2) return oracle_deriv_kernel(1, tk, tj, l);
  if (kind == 5) return oracle_deriv_kernel(4, tk, tj, l);
  if (kind == 7) return oracle_deriv_kernel(6, tk, tj, l);
  const double d = tj - tk, e = exp(-(d * d / (2.0 * l * l)));
  const double l2 = l * l, l4 = l2 * l2, l6 = l4 * l2, l8 = l4 * l4;
  switch (kind) {
    case 0: return e;
    case 1: return (e * d) / l2;
    case 3: return e / l2 - (e * d * d) / l4;
    case 4: return -(e / l2) + (e * d * d) / l4;
    case 6: return (3.0 * e * d) / l4 - (e * d * d * d) / l6;
    case 8: return (3.0 * e) / l4 - (6.0 * e * d * d) / l6 + (e * d * d * d * d) / l8;
    default: return NAN;
  }
}

void oracle_outer_kernel(int kind, int n, const double *tj, int m, const double *tk, double l,
                         double amp2, double *K) {
  for (int j = 0; j < m; j++)
    for (int i = 0; i < n; i++) K[IDX(i, j, n)] = amp2 * oracle_deriv_kernel(kind, tj[i], tk[j], l);
}
