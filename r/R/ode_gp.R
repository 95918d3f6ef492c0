# Replacement for the older R/ode_gp.R of bbbales2/gp: the explicit-solve flavour that returns list(mn, Kn)
# (R/tests.R:37,90 reads $mn / $Kn).  Needs r/R/gpb200.R and r/R/kernels.R sourced first.

# R/ode_gp.R:1-14 -- mn = K (K + sigma^2 I)^-1 Xn ; Kn = K - K (K + sigma^2 I)^-1 K
p_Xn <- function(tn, Xn, phi_n, sigma_n) {
  K <- QQ(tn, tn, phi_n)
  m <- gp_condition(K, K, K, Xn, sigma_n^2, 0)
  list(mn = matrix(m$mu, ncol = 1), Kn = m$cov)
}

# R/ode_gp.R:19-32 -- mn = RQ (QQ + sigma^2 I)^-1 Xn ; Kn = RR - RQ (QQ + sigma^2 I)^-1 QR.  One Cholesky-based
# conditioning call on the GPU instead of two LU solves with an N x N right-hand side.
p_dotXn <- function(tn, Xn, phi_n, sigma_n) {
  K <- QQ(tn, tn, phi_n)
  RQm <- t(QR(tn, tn, phi_n))
  m <- gp_condition(K, RQm, RR(tn, tn, phi_n), Xn, sigma_n^2, 0)
  list(mn = matrix(m$mu, ncol = 1), Kn = m$cov)
}

# R/ode_gp.R:35-37 -- empty in the reference
p_dotX <- function(X, phi, sigma_sq) {

}

# R/ode_gp.R:42-103 -- the older sequential sampler.  The reference text cannot run as written (`qr.` :55,
# `QQard(xs, xs)` without theta :74, `rnorm(mu_s2, sigma_s2)` :93); its intent is the two-block conditional that
# R/ode_gp_library.R:43-93 states in full, so this is that sampler with the older return names (mu, sigma, dotxs).
create_p_dotXnS <- function(Xn_list, mn, Kn, theta) {
  env <- new.env()
  sys.source(file.path(Sys.getenv("GPB200_HOME", "."), "r", "R", "ode_gp_library.R"), envir = env)
  Xl <- if (is.list(Xn_list)) Xn_list else list(Xn_list)
  step <- env$create_p_dotXnS(Xl, mn, Kn, theta)
  function(xs) {
    r <- step(xs)
    list(mu = r$mu, sigma = r$sigma, dotxs = r$dot_xs)
  }
}
