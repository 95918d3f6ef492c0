# Replacement for R/ode_gp_library.R of bbbales2/gp (same names, same formals, same return values).
# Needs r/R/gpb200.R and r/R/kernels.R sourced first.  condMVN is provided by gpb200.R (the reference takes it
# from the condMVNorm package, R/ode_gp_library.R:1).

# UU / UD / DD are called by the reference (R/ode_gp_library.R:8-11,29-30) but defined nowhere in its tree; by
# position they are the value / value-derivative / derivative-derivative kernels of R/kernels.R, as R/ode_gp.R:5-8,
# 23-26 spells out (SURVEY Appendix A.2).
UU <- function(x, y, phi) QQ(x, y, phi)
UD <- function(x, y, phi) QR(x, y, phi)
DD <- function(x, y, phi) RR(x, y, phi)

# R/ode_gp_library.R:3-18
p_Xn <- function(tn, Xn, phi_n, sigma_n) {
  N <- length(Xn)
  K_XX <- UU(tn, tn, phi_n)
  K <- rbind(cbind(K_XX + sigma_n^2 * diag(N), t(K_XX)), cbind(t(K_XX), K_XX)) + 1e-6 * diag(2 * N)
  condMVN(rep(0, 2 * N), K, (N + 1):(2 * N), 1:N, Xn)
}

# R/ode_gp_library.R:23-33 -- the 2N x 2N joint covariance is assembled by one Gram kernel (blocks QQ, QR; RQ, RR with
# the R/kernels.R:31 quirk, sigma_n^2 on the first block, 1e-6 on the whole diagonal) and conditioned by the Cholesky
p_dotXn <- function(tn, Xn, phi_n, sigma_n) {
  N <- length(Xn)
  K <- .Call("gp_gram_deriv", tn, phi_n[[1]], phi_n[[2]], 2L, c(sigma_n, 0), 1e-6, 1L)
  condMVN(rep(0, 2 * N), K, (N + 1):(2 * N), 1:N, Xn)
}

# R/ode_gp_library.R:36-38 -- an empty stub in the reference; kept so that code referring to it still parses
p_dotX <- function(X, phi, sigma_sq) {

}

# R/ode_gp_library.R:43-93 -- sequential conditional sampler of the derivative at new states.
# Same closure protocol: p <- create_p_dotXnS(Xn_list, mn, Kn, theta); p(xs) -> list(mu, sigma, dot_xs).
# Differences in HOW, not WHAT: K_XX + 1e-6 I is Cholesky-factored on the GPU once instead of qr() (:55-57); each
# call solves only for the NEW cross-covariance column and extends the lower Cholesky factor of the star points'
# joint covariance by one row (a bordered update, O(i^2)) instead of rebuilding and re-solving the whole i x i system
# (:75-82); the conditional mean / variance of the new point given the earlier draws then fall out of that row.
create_p_dotXnS <- function(Xn_list, mn, Kn, theta) {
  X <- do.call(cbind, Xn_list)
  N <- nrow(X)
  D <- ncol(X)
  L_XX <- gp_chol(QQard(X, X, theta) + 1e-6 * diag(N))
  K_XX_1_mn <- gp_chol_solve(L_XX, mn)
  K_XX_1_Kn <- gp_chol_solve(L_XX, Kn)

  i <- 1
  Xs <- matrix(nrow = 0, ncol = D)        # star points seen so far
  S <- matrix(nrow = N, ncol = 0)         # K_XX^-1 K_XXs, one column per star point
  A <- matrix(nrow = 0, ncol = N)         # K_XsX
  Lc <- matrix(nrow = 0, ncol = 0)        # lower Cholesky factor of the star points' joint covariance
  w <- numeric(0)                         # Lc^-1 (dot_Xs - m)

  p_dotXnS <- function(xs_vec) {
    xs <- matrix(xs_vec, nrow = 1)
    a <- QQard(xs, X, theta)                                  # 1 x N
    s <- gp_chol_solve(L_XX, as.numeric(a))                   # new column of solve(K_XX, t(K_XsX))
    A <<- rbind(A, a)
    S <<- cbind(S, s)
    m_new <- as.numeric(a %*% K_XX_1_mn)
    kss <- as.numeric(QQard(Xs, xs, theta))                   # cross-covariances with the earlier star points
    # new row of K = K_XsXs - K_XsX K_XX^-1 K_XXs + K_XsX K_XX^-1 Kn K_XX^-1 K_XXs, symmetrised, + 1e-6 on the diagonal
    left <- as.numeric(a %*% S) - as.numeric((a %*% K_XX_1_Kn) %*% S)
    right <- as.numeric(A %*% s) - as.numeric((A %*% K_XX_1_Kn) %*% s)
    krow <- c(kss, as.numeric(QQard(xs, xs, theta))) - (left + right) / 2
    krow[i] <- krow[i] + 1e-6
    if (i == 1) {
      l_row <- numeric(0)
      cond_mean <- m_new
      cond_var <- krow[1]
    } else {
      l_row <- forwardsolve(Lc, krow[1:(i - 1)])              # bordered Cholesky: new row of the factor
      cond_mean <- m_new + sum(l_row * w)
      cond_var <- krow[i] - sum(l_row^2)
    }
    # bug-compatible with R/ode_gp_library.R:84, which hands a VARIANCE to rnorm's sd argument (SURVEY Appendix A.3)
    dot_xs <- rnorm(1, cond_mean, cond_var)
    d <- sqrt(cond_var)
    Lc <<- rbind(cbind(Lc, matrix(0, nrow = i - 1, ncol = 1)), c(l_row, d))
    w <<- c(w, (dot_xs - cond_mean) / d)
    i <<- i + 1
    Xs <<- rbind(Xs, xs)
    list(mu = cond_mean, sigma = cond_var, dot_xs = dot_xs)
  }
}
