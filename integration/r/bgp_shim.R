# bgp_shim.R — R-side glue over bgp_rcall.c (UNTESTED here: no R in the build image).
# Drop-in for the three third-party calls of get_result_by_method()/model_fit()
# (/root/reference/R/02_model_fit.R:276-284 and :687-689).  Everything above (formula parsing, S4 term
# objects, tmbdat construction) and everything below in R/03_post_fit.R keeps working on the returned lists.

# ff <- TMB::MakeADFun(data = tmbdat, parameters = tmbparams, random = "W", DLL = "BayesGP", silent = TRUE)
bgp_MakeADFun <- function(tmbdat, tmbparams, device = 0L) {
  dense <- function(l) lapply(l, as.matrix)          # undo the dgTMatrix conversion (R/01_utility.R:484-488)
  tmbdat$X <- dense(tmbdat$X); tmbdat$B <- dense(tmbdat$B); tmbdat$P <- dense(tmbdat$P); tmbdat$Xf <- dense(tmbdat$Xf)
  h <- .Call("bgpR_model", tmbdat, as.integer(device))
  env <- new.env()
  env$random <- seq_along(tmbparams$W)
  env$last.par <- c(tmbparams$W, tmbparams$theta)
  ev <- function(theta, grad = FALSE, hess = FALSE) {
    r <- .Call("bgpR_eval", h, as.double(theta), grad, TRUE, hess)
    if (!is.null(r[[3]])) env$last.par <- c(r[[3]], theta)
    if (hess) env$last.H <- r[[4]]
    r
  }
  env$spHess <- function(par = env$last.par, random = TRUE) {
    theta <- par[-env$random]
    Matrix::Matrix(ev(theta, hess = TRUE)[[4]], sparse = TRUE)
  }
  ff <- list(par = tmbparams$theta, env = env, handle = h,
             fn = function(theta) ev(theta)[[1]][1],
             gr = function(theta) matrix(ev(theta, grad = TRUE)[[2]], nrow = 1))
  ff$he <- function(w) numDeriv::jacobian(ff$gr, w)   # R/02_model_fit.R:283 (kept for callers that use it)
  ff
}

# mod <- aghq::marginal_laplace_tmb(ff, k = aghq_k, startingvalue = rep(0, S))
bgp_marginal_laplace <- function(ff, k, startingvalue) {
  fit <- .Call("bgpR_fit", ff$handle, as.integer(k), as.double(startingvalue))
  g <- .Call("bgpR_fit_get", fit)
  S <- length(g[[1]]); K <- length(g[[5]])
  nw <- as.data.frame(g[[4]]); names(nw) <- paste0("theta", seq_len(S))
  nw$weights <- g[[5]]; nw$logpost <- g[[6]]; nw$logpost_normalized <- g[[7]]
  mh <- nw[, seq_len(S), drop = FALSE]
  mh$mode <- lapply(seq_len(K), function(j) g[[9]][, j])
  mh$H <- lapply(seq_len(K), function(j) Matrix::Matrix(g[[10]][, , j], sparse = TRUE))
  marg <- lapply(seq_len(S), function(j) {
    d <- as.data.frame(g[[11]][[j]]); names(d) <- c(paste0("theta", j), "logmargpost", "w"); d
  })
  structure(list(
    normalized_posterior = list(nodesandweights = nw, grid = list(level = rep(k, S)), lognormconst = g[[8]]),
    marginals = marg,
    optresults = list(ff = ff, mode = g[[1]], hessian = g[[2]], convergence = g[[3]]),
    modesandhessians = mh, control = list(method = "BFGS", negate = TRUE), transformation = NULL, handle = fit),
    class = c("marginallaplace", "aghq"))
}

# samps <- aghq::sample_marginal(mod, M)   (R's RNG draws the node ids and Z; the library does the algebra)
bgp_sample_marginal <- function(mod, M) {
  nw <- mod$normalized_posterior$nodesandweights
  lambda <- nw$weights * exp(nw$logpost_normalized)
  idx <- sample.int(length(lambda), M, replace = TRUE, prob = lambda) - 1L
  p <- length(mod$modesandhessians$mode[[1]])
  Z <- matrix(stats::rnorm(p * M), p, M)
  S <- sum(grepl("^theta", names(nw)))
  list(samps = .Call("bgpR_sample", mod$handle, Z, as.integer(idx)),
       theta = nw[idx + 1L, seq_len(S), drop = FALSE])
}
