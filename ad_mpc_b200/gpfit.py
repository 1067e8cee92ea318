"""Host-side mirror of `CustomGPRegression.fit` (reference: model_fitting/gp.py:325-369) with the dense linear algebra
on the GPU: the NLL the reference minimises (gp.py:305-311) and the final K^-1 y (gp.py:361-363) are evaluated by
`admpc_gp_fit`; the hyper-parameter search is the same L-BFGS-B in log space with the same bounds (gp.py:336-338).
The result is the dict `BatchSolver.set_gp` consumes, so an online GP refresh is fit -> set_gp (-> NCCL broadcast)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def nll_alpha(X, y, ell, sigma_f, sigma_n, device=0, want_alpha=True):
    """(nll, alpha, device ms) for one output dimension; y with its mean already removed."""
    L = _lib.load()
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    ell = np.ascontiguousarray(np.broadcast_to(np.asarray(ell, dtype=np.float64).reshape(-1), (X.shape[1],)))
    alpha = np.empty(X.shape[0]) if want_alpha else None
    nll, ms = C.c_double(), C.c_float()
    check(L.admpc_gp_fit(int(device), X.shape[0], X.shape[1], _dp(X), _dp(y), _dp(ell), float(sigma_f), float(sigma_n),
                         _dp(alpha) if want_alpha else None, C.byref(nll), C.byref(ms)), "admpc_gp_fit")
    return nll.value, alpha, ms.value


def fit(X, y, ell0=None, sigma_f0=1.0, sigma_n0=1e-2, device=0, optimise=True):
    """Fit one output dimension. Returns dict(X, alpha, ell, sigma_f, sigma_n, y_mean, nll)."""
    from scipy.optimize import minimize
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    dz = X.shape[1]
    y_mean = float(np.mean(y))
    yc = y - y_mean                                                  # gp.py:343
    theta0 = np.log(np.r_[np.ones(dz) if ell0 is None else np.asarray(ell0, dtype=np.float64), sigma_f0, sigma_n0])
    bounds = [(np.log(1e-5), np.log(1e1))] * (dz + 1) + [(np.log(1e-8), np.log(1e0))]          # gp.py:336-338

    def f(theta):
        p = np.exp(theta)
        try:
            return nll_alpha(X, yc, p[:dz], p[dz], p[dz + 1], device=device, want_alpha=False)[0]
        except _lib.AdmpcError:
            return np.inf                                            # LinAlgError branch, gp.py:322-323

    theta = theta0
    if optimise:
        res = minimize(f, x0=theta0, bounds=bounds, method="L-BFGS-B")
        theta = res.x
    p = np.exp(theta)
    nll, alpha, _ = nll_alpha(X, yc, p[:dz], p[dz], p[dz + 1], device=device)
    return dict(X=X, y=yc, alpha=alpha, ell=p[:dz].copy(), sigma_f=float(p[dz]), sigma_n=float(p[dz + 1]), y_mean=y_mean, nll=nll)


def predict(model, x_test, y=None, device=0, return_cov=False, full_cov=False):
    """Mirror of `CustomGPRegression.predict(x_test, return_std/return_cov)` (gp.py:402-441) for one fitted output
    (dict from `fit`, plus the mean-removed training targets `y` if the dict does not carry them as model["y"]).
    Returns mu, or (mu, var) / (mu, var, cov[n,n])."""
    L = _lib.load()
    X = np.ascontiguousarray(model["X"], dtype=np.float64)
    yc = np.ascontiguousarray(model["y"] if y is None else y, dtype=np.float64).reshape(-1)
    xt = np.ascontiguousarray(np.atleast_2d(np.asarray(x_test, dtype=np.float64)))
    ell = np.ascontiguousarray(np.broadcast_to(np.asarray(model["ell"], dtype=np.float64).reshape(-1), (X.shape[1],)))
    n = xt.shape[0]
    mu, var = np.empty(n), np.empty(n)
    cov = np.empty((n, n)) if full_cov else None
    check(L.admpc_gp_predict(int(device), X.shape[0], X.shape[1], _dp(X), _dp(yc), _dp(ell), float(model["sigma_f"]),
                             float(model["sigma_n"]), float(model["y_mean"]), n, _dp(xt), _dp(mu), _dp(var),
                             _dp(cov) if full_cov else None), "admpc_gp_predict")
    if full_cov:
        return mu, var, cov
    return (mu, var) if return_cov else mu


def stack_models(models, feat=(3, 4, 5, 6), rows=(4, 5)):
    """Per-output fits -> the multi-output dict of BatchSolver.set_gp (all outputs must share M and dz)."""
    return dict(X=np.stack([m["X"] for m in models]), alpha=np.stack([m["alpha"] for m in models]),
                ell=np.stack([m["ell"] for m in models]), sigma_f=np.array([m["sigma_f"] for m in models]),
                y_mean=np.array([m["y_mean"] for m in models]), feat=tuple(feat), rows=tuple(rows))
