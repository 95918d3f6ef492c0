# Replacement for derivative_kernels.R:39-73 of bbbales2/gp: the nine unit-amplitude covariance functions of a
# squared-exponential GP and its first (R) and second (T) derivatives, under the reference's own names and
# formals (tj, tk, l).  First letter goes with tj.  Vectorised over tj / tk like the originals (so that
# outer(ti, ti, FUN = function(a, b) RQ(a, b, l)) of pendulum_fit.R:238-240 keeps working); each call is one
# GPU kernel over all elements.  Needs r/R/gpb200.R sourced first.
QQ <- function(tj, tk, l) gp_elementwise("QQ", tj, tk, l)
QR <- function(tj, tk, l) gp_elementwise("QR", tj, tk, l)
RQ <- function(tj, tk, l) gp_elementwise("RQ", tj, tk, l)
RR <- function(tj, tk, l) gp_elementwise("RR", tj, tk, l)
QT <- function(tj, tk, l) gp_elementwise("QT", tj, tk, l)
TQ <- function(tj, tk, l) gp_elementwise("TQ", tj, tk, l)
RT <- function(tj, tk, l) gp_elementwise("RT", tj, tk, l)
TR <- function(tj, tk, l) gp_elementwise("TR", tj, tk, l)
TT <- function(tj, tk, l) gp_elementwise("TT", tj, tk, l)
