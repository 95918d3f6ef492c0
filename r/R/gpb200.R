# Drop-in R definitions backed by libgpb200.so (through r/shim.c).  Source this file INSTEAD of
#   sourceCpp("covariance.cpp")            (gpc_sigma.R:7)
#   source("R/kernels.R"), source("derivative_kernels.R"), source("R/ode_gp_library.R")
# Formals are identical to the reference's; each matrix is built by ONE .Call instead of outer().
dyn.load(file.path(Sys.getenv("GPB200_HOME", "."), "r", "gpb200_r.so"))

.gp_kind <- c(QQ = 0L, QR = 1L, RQ = 2L, RR = 3L, QT = 4L, TQ = 5L, RT = 6L, TR = 7L, TT = 8L, RR_QUIRK = 9L)

# ---- covariance.cpp -----------------------------------------------------------------------------
rbf_cov_chol <- function(x1, l_) .Call("gp_rbf_cov_chol", as.double(x1), as.double(l_))
approx_L <- function(l, lp, Ls, dLdls) .Call("gp_approx_L", as.double(l), as.double(lp), Ls, dLdls)
# all tables of a length-scale grid in one GPU call: tabs <- rbf_cov_chol_grid(x, lp); tabs$Ls, tabs$dLdls
rbf_cov_chol_grid <- function(x1, lp) .Call("gp_rbf_cov_chol_grid", as.double(x1), as.double(lp))
# eigen-basis factor of models/westbrook.stan:2-30 (named approx_L there too; bH in spectral_test.R:6)
bH <- function(M, scale, x, sigma, l) .Call("gp_approx_L_basis", as.integer(M), as.double(scale), as.double(x), as.double(sigma), as.double(l))
# latent models: Cholesky and its tangent (wrt = 0 alpha, 1 rho)
se_chol_tangent <- function(x, alpha, rho, diag_add, wrt) .Call("gp_se_chol_tangent", as.double(x), as.double(alpha), as.double(rho), as.double(diag_add), as.integer(wrt))

# ---- derivative_kernels.R:39-73 (element-wise closures, vectorised like the originals) ------------
.gp_elem <- function(kind) function(tj, tk, l) {
  n <- max(length(tj), length(tk))
  .Call("gp_kernel_eval", .gp_kind[[kind]], rep_len(as.double(tj), n), rep_len(as.double(tk), n), 1.0, as.double(l))
}
dk_QQ <- .gp_elem("QQ"); dk_QR <- .gp_elem("QR"); dk_RQ <- .gp_elem("RQ"); dk_RR <- .gp_elem("RR")
dk_QT <- .gp_elem("QT"); dk_TQ <- .gp_elem("TQ"); dk_RT <- .gp_elem("RT"); dk_TR <- .gp_elem("TR"); dk_TT <- .gp_elem("TT")
# outer(ti, ti, FUN = kern_fixed_l(RQ, l)) of pendulum_fit.R:238-240 in one call:
gp_outer <- function(kind, tj, tk, l, amp2 = 1.0) .Call("gp_gram_outer", .gp_kind[[kind]], as.double(tj), as.double(tk), as.double(amp2), as.double(l))

# ---- R/kernels.R:19-32 (phi = c(alpha, rho); the derivative_kernels.R names collide with these,
#      exactly as they do in the reference -- source the flavour you need last) ----------------------
QQ <- function(x, y, phi) gp_outer("QQ", x, y, phi[[2]], phi[[1]]^2)
QR <- function(x, y, phi) gp_outer("QR", x, y, phi[[2]], phi[[1]]^2)
RR <- function(x, y, phi) gp_outer("RR_QUIRK", x, y, phi[[2]], phi[[1]]^2)   # bug-compatible with R/kernels.R:31
QQard <- function(X, Y, phi) .Call("gp_gram_ard", X, Y, as.double(phi[[1]]), as.double(unlist(phi[[2]])))

# ---- R/ode_gp_library.R ---------------------------------------------------------------------------
p_dotXn <- function(tn, Xn, phi_n, sigma_n) {
  N <- length(Xn)
  K <- .Call("gp_gram_deriv", as.double(tn), as.double(phi_n[[1]]), as.double(phi_n[[2]]), 2L,
             c(as.double(sigma_n), 0), 1e-6, 1L)
  .Call("gp_cond_mvn", rep(0, 2 * N), K, as.integer(N), as.double(Xn))
}
p_Xn <- function(tn, Xn, phi_n, sigma_n) {
  N <- length(Xn)
  UU <- QQ(tn, tn, phi_n)
  K <- rbind(cbind(UU + sigma_n^2 * diag(N), t(UU)), cbind(t(UU), UU)) + 1e-6 * diag(2 * N)
  .Call("gp_cond_mvn", rep(0, 2 * N), K, as.integer(N), as.double(Xn))
}

# ---- pendulum_fit.R:227-255 -----------------------------------------------------------------------
sample_derivs <- function(params, ynoise, ti, seed = NULL) {
  l <- params[1]; a <- params[2]; sy <- params[3]
  K <- gp_outer("QQ", ti, ti, l, a^2); KsK <- gp_outer("RQ", ti, ti, l, a^2); KsKs <- gp_outer("RR", ti, ti, l, a^2)
  m <- .Call("gp_condition", K, KsK, KsKs, as.double(ynoise), sy^2, 1e-8)
  if (!is.null(seed)) return(as.numeric(.Call("gp_mvrnorm", 1L, m$mu, m$cov, as.double(seed))))  # device RNG
  L <- .Call("gp_potrf", m$cov)
  as.numeric(m$mu + L %*% rnorm(length(m$mu)))
}
mvrnorm <- function(n = 1, mu, Sigma, seed = sample.int(.Machine$integer.max, 1)) {   # MASS::mvrnorm signature + seed
  out <- .Call("gp_mvrnorm", as.integer(n), as.double(mu), Sigma, as.double(seed))
  if (n == 1) drop(out) else out
}

# ---- batched LML + gradient over hyper-parameter draws (the mclapply axis, pendulum_fit.R:259-268) -
gp_lml_grad_draws <- function(x, y, theta, jitter = 0) .Call("gp_lml_grad_draws", as.double(x), as.double(y), theta, as.double(jitter))

# ---- GP observed through derivatives (the Stan model inside gpderivs.py:25-133; design_notes.Rmd:25-46) ----
# theta: (2 + nblocks) x B matrix, columns (alpha, rho, noise_1..noise_nblocks); y = c(y, yp, ypp) stacked
gp_lml_grad_deriv_draws <- function(t, y, theta, order0 = 0L, jitter = 0)
  .Call("gp_lml_grad_deriv_draws", as.double(t), as.double(y), theta, as.integer(order0), as.double(jitter))
