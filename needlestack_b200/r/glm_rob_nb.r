# Drop-in replacement for bin/glm_rob_nb.r: same name, same argument list, same result list; the body
# packs the unit and calls the B200 library through the .Call shim (ns_rshim.c).
# NOT RUN IN THIS IMAGE (R is not installed here).  Arguments that the reference ignores are ignored
# (maxmu, overwritten at bin/glm_rob_nb.r:20; c.by.beta, n_xout); unsupported options stop() with the
# reference's own messages (bin/glm_rob_nb.r:99,226).

.ns_env <- new.env()
.ns_ctx <- function(device = as.integer(Sys.getenv("NEEDLESTACK_B200_DEVICE", "0"))) {
  if (is.null(.ns_env$ctx)) {
    dyn.load(Sys.getenv("NEEDLESTACK_B200_SHIM", "ns_rshim.so"))
    .ns_env$ctx <- .Call("C_ns_create", device)
  }
  .ns_env$ctx
}

glmrob.nb <- function(y,x,bounding.func='T/T',c.tukey.beta=5,c.tukey.sig=3,c.by.beta=4,weights.on.x='none',
                      minsig=1e-3,maxsig=10,minmu=1e-10,maxmu=1e5,maxit=30,maxit.sig=50,sig.prec=1e-8,tol=1e-6,
                      n_ai.sig.tukey=100,n_xout=10^4,min_coverage=1,min_reads=1,min_af=0,size_min=10,extra_rob=TRUE,
                      min_af_extra_rob=0.2,min_prop_extra_rob=0.1,max_prop_extra_rob=0.5,...){
  if (!(weights.on.x %in% c('none','hard'))) stop('Only "hard" and "none" are implemented for weights.on.x.')
  if (weights.on.x == 'hard') stop('weights.on.x="hard" (MASS::cov.rob) is not available on the GPU path')
  if (bounding.func != 'T/T') stop('Available bounding.func is "T/T"')
  if (length(y) != length(x)) stop('length(y) does not match dim(X)[1].')
  if (any(y != round(y)) || any(x != round(x))) stop('needlestack-b200: counts must be integers')
  r <- .Call("C_ns_glmrob_nb", .ns_ctx(), as.integer(y), as.integer(x),
             list(c_tukey_beta=c.tukey.beta, c_tukey_sig=c.tukey.sig, minsig=minsig, maxsig=maxsig, minmu=minmu,
                  maxit=maxit, maxit_sig=maxit.sig, sig_prec=sig.prec, tol=tol, n_ai_sig_tukey=n_ai.sig.tukey,
                  min_coverage=min_coverage, min_reads=min_reads, min_af=min_af, size_min=size_min,
                  extra_rob=as.integer(extra_rob), min_af_extra_rob=min_af_extra_rob,
                  min_prop_extra_rob=min_prop_extra_rob, max_prop_extra_rob=max_prop_extra_rob))
  if (bitwAnd(r$flags, 0x20L) != 0) stop("integrate() failed inside the IRLS loop")  # NS_F_QUAD_ERROR: as the reference
  list("coverage"=x, "ma_count"=y, "coef"=c(sigma=r$sigma, slope=r$slope), "pvalues"=r$pvalues,
       "qvalues"=r$qvalues, "GQ"=r$GQ, "extra_rob"=r$extra_rob)
}
