#!/usr/bin/env Rscript
# Cross-check of the golden fixtures against the REFERENCE glmrob.nb() - for any machine that has R (this image does
# not: parity with R itself is unpinned, see DESIGN.md section 2).
#   python tests/r/export_cases.py > cases.tsv
#   Rscript tests/r/crosscheck.R /path/to/needlestack/bin/glm_rob_nb.r cases.tsv
# Prints, per case, the largest relative difference of sigma / slope and the largest absolute difference of GQ between
# the reference and the fixture (the oracle's output, which the CUDA path matches to 1e-6), and fails above tolerance.
args <- commandArgs(TRUE)
source(args[1])
lines <- readLines(args[2])
worst <- list()
for (ln in lines) {
  f <- strsplit(ln, "\t", fixed = TRUE)[[1]]
  n <- as.integer(f[3])
  sigma <- as.numeric(f[4]); slope <- as.numeric(f[5])
  y <- as.numeric(f[6:(5 + n)]); x <- as.numeric(f[(6 + n):(5 + 2 * n)]); gq <- as.numeric(f[(6 + 2 * n):(5 + 3 * n)])
  # needlestack.r's call (bin/needlestack.r:336-337) with its default options (:64-92)
  r <- glmrob.nb(x = x, y = y, min_coverage = 30, min_reads = 3, min_af = 0, extra_rob = FALSE,
                 min_af_extra_rob = 0.2, min_prop_extra_rob = 0.1, max_prop_extra_rob = 0.5)
  d <- c(abs(r$coef[["sigma"]] - sigma) / sigma, abs(r$coef[["slope"]] - slope) / slope,
         max(abs(r$GQ - gq) / pmax(1, abs(gq))))
  w <- worst[[f[1]]]
  worst[[f[1]]] <- if (is.null(w)) d else pmax(w, d)
}
bad <- FALSE
for (k in names(worst)) {
  cat(sprintf("%-12s max rel d(sigma) %.3g  d(slope) %.3g  d(GQ) %.3g\n", k, worst[[k]][1], worst[[k]][2], worst[[k]][3]))
  # sigma: the reference stops Brent at an absolute 1e-8, so small sigmas may differ by 1e-5 relative
  if (worst[[k]][2] > 1e-6 || worst[[k]][3] > 1e-5) bad <- TRUE
}
if (bad) quit(status = 1)
