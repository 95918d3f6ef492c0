#!/bin/bash
# Builds and smoke-tests the R shim wherever R exists (not in the build container).
set -e
cd "$(dirname "$0")/.."
python gp_b200/build.py
PKG_CPPFLAGS="-Iinclude" PKG_LIBS="-Lgp_b200/lib -lgpb200 -Wl,-rpath,$PWD/gp_b200/lib" R CMD SHLIB r/shim.c -o r/gpb200_r.so
Rscript -e 'source("r/R/gpb200.R"); x <- seq(0, 10, length = 100); o <- rbf_cov_chol(x, 1.0); stopifnot(all(dim(o$L) == c(100, 100)), max(abs(o$L %*% t(o$L) - (exp(-outer(x, x, "-")^2 / 2) + 1e-10 * diag(100)))) < 1e-10); cat("R shim ok\n")'
