# Replacement for R/kernels.R of bbbales2/gp (same names, same formals).  Needs r/R/gpb200.R sourced first.

# R/kernels.R:2-17 -- the generic pairwise builder for a user-supplied K(x, y, phi).  An arbitrary R closure
# cannot run on the GPU; these three keep the reference's behaviour on the host so that code written against
# them still works.  The kernels the reference actually builds with it (QQard below) have GPU versions.
mat_to_obs_list <- function(X) lapply(seq_len(nrow(X)), function(i) X[i, ])
obs_list_outer <- function(X, Y, K) outer(X, Y, function(a, b) vapply(seq_along(a), function(i) K(a[[i]], b[[i]]), numeric(1)))
create_kernel_function <- function(K) function(X, Y, phi) obs_list_outer(mat_to_obs_list(X), mat_to_obs_list(Y), function(x, y) K(x, y, phi))

# R/kernels.R:19 -- ARD squared exponential, phi = list(alpha, rho[D] or scalar): one GPU call for the matrix
QQard <- function(X, Y, phi) {
  X <- if (is.matrix(X)) X else matrix(X, ncol = 1)
  Y <- if (is.matrix(Y)) Y else matrix(Y, ncol = 1)
  .Call("gp_gram_ard", X, Y, phi[[1]], unlist(phi[[2]]))
}

# R/kernels.R:22-32 -- 1-D kernels, phi = c(alpha, rho)
QQ <- function(x, y, phi) gp_outer("QQ", x, y, phi[[2]], phi[[1]]^2)
QR <- function(x, y, phi) gp_outer("QR", x, y, phi[[2]], phi[[1]]^2)
# bug-compatible with R/kernels.R:31, where phi[[1]]^2 multiplies only the first term (SURVEY Appendix A.1);
# gp_outer("RR", ...) is the mathematically consistent kernel
RR <- function(x, y, phi) gp_outer("RR_QUIRK", x, y, phi[[2]], phi[[1]]^2)
