## R-side dispatch for the CUDA backend -- SOURCE ONLY here (no R in the build image).
##
## Add to the reference's R/ directory.  With options(mnem.backend = "cuda") the EM phase of mnem() is one call
## into libmnemcu.so; everything before it (modData, learnk, initComps / sampleRndNetwork with R's RNG) and after
## it (best-start selection is returned by the library but re-checked here, result assembly, getIC, plotting) is
## the reference's own code, untouched.

mnem_cuda_enabled <- function() identical(getOption("mnem.backend", "r"), "cuda")

## replaces the `lapply(seq_len(starts), do_inits)` of R/mnems.r:2166-2170 for inference = "em", search = "greedy",
## marginal = FALSE; `init` is mnem()'s phi list (R/mnems.r:1790-1793).  Multi-knock-down columns ("1_3") or an
## explicit Rho take the Rho path (R/mnems_low.r:613-616, 896-904).
mnem_cuda_do_inits <- function(data, init, complete, logtype, device = 0, Rho = NULL, affinity = 0, mw = NULL,
                               nullcomp = FALSE, theta = NULL) {
    if (is.null(Rho) && any(grepl("_", colnames(data)))) Rho <- getRho(data)
    if (is.null(Rho)) {
        pert <- as.integer(colnames(data))                   # after modData(): "1".."S"
        S <- length(unique(pert))
        ctx <- mnemcu_context(data, pert, S, device)
    } else {
        S <- nrow(Rho)
        ctx <- mnemcu_context_rho(data, Rho, device)
    }
    if (nullcomp) {                                          # res1[[k + 1]]$adj <- adj * 0  (R/mnems.r:1894-1898)
        init <- lapply(init, function(g) c(g, list(g[[1]] * 0)))
        if (!is.null(theta)) theta <- lapply(theta, function(t) c(t, list(t[[1]] * 0L)))
    }
    out <- mnemcu_em(ctx, init, complete, logtype, 0L, 256L, as.integer(affinity), mw, nullcomp, theta)
    for (s in seq_along(out$limits)) {                       # names the rest of mnem() expects (R/mnems.r:2153-2163)
        out$limits[[s]]$k <- length(init[[s]])
        out$limits[[s]]$subtopo <- NULL
        for (i in seq_along(out$limits[[s]]$res)) {
            dimnames(out$limits[[s]]$res[[i]]$adj) <- list(seq_len(S), seq_len(S))
        }
    }
    out$limits
}

## In mnem() (R/mnems.r:2166): 
##   if (mnem_cuda_enabled() && is.null(parallel)) {
##       limits <- mnem_cuda_do_inits(data, init, complete, logtype, affinity = affinity, mw = mw,
##                                    nullcomp = nullcomp, theta = theta)
##   } else if (!is.null(parallel)) { limits <- sfLapply(seq_len(starts), do_inits) } else { ... }
## snowfall (`parallel = n`) and the CUDA backend must not be combined: one CUDA context per process.

## In getAffinity() (R/mnems.r:1390), soft path with norm = TRUE:
##   if (mnem_cuda_enabled()) return(mnemcu_getAffinity(x, if (is.null(mw)) rep(1, nrow(x))/nrow(x) else mw,
##                                                      affinity, complete, logtype))
## In scoreAdj() (R/mnems.r:416) for method = "llr", weights = NULL, D already collapsed:
##   if (mnem_cuda_enabled()) return(mnemcu_scoreAdj(D, adj, trans.close, marginal, logtype))
## In nem() (R/mnems.r:52) after doMean (line 107), search = "greedy", runs = 1:
##   if (mnem_cuda_enabled()) { fit <- mnemcu_nem(D, start, marginal, logtype); ... assemble list(adj, score, ...) }
