"""CPU ORACLE (test infrastructure only) for GP fitting / model update.

numpy restatement of the numeric core of `CustomGPRegression` (reference: model_fitting/gp.py): the training kernel
matrix of the two-argument kernel call (:103-105, sigma_f * exp(-0.5 * sqeuclidean(x/l, x/l)), diagonal = sigma_f),
the negative log likelihood (:283-289 / :305-311) and K^-1 y (:361-363).
Pinned against outputs of the reference module itself: tests/golden/gp_reference.npz (tests/golden/make_gp_golden.py).
"""
import numpy as np


def train_kernel(X, ell, sigma_f, sigma_n):
    Xs = X / ell
    d2 = ((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(axis=2)
    return sigma_f * np.exp(-0.5 * d2) + sigma_n ** 2 * np.eye(X.shape[0])


def nll(X, y, ell, sigma_f, sigma_n):
    """gp.py:305-311 with theta already exponentiated."""
    K = train_kernel(X, ell, sigma_f, sigma_n)
    L = np.linalg.cholesky(K)
    a = np.linalg.solve(L.T, np.linalg.solve(L, y))
    return float(np.sum(np.log(np.diagonal(L))) + 0.5 * y @ a + 0.5 * X.shape[0] * np.log(2 * np.pi))


def alpha(X, y, ell, sigma_f, sigma_n):
    """K^-1 y (gp.py:361-363; the reference forms inv(K) explicitly, here by Cholesky)."""
    K = train_kernel(X, ell, sigma_f, sigma_n)
    L = np.linalg.cholesky(K)
    return np.linalg.solve(L.T, np.linalg.solve(L, y))


def predict(X, y, ell, sigma_f, sigma_n, y_mean, xt):
    """Posterior mean and variance of `CustomGPRegression.predict(x_test, return_cov=True)` (gp.py:402-441):
    k_s = k(x_test, X); mu = k_s K^-1 y + y_mean; cov = k(x_test, x_test) + 1e-8 I - k_s K^-1 k_s^T, diagonal returned.
    K^-1 is formed explicitly like the reference does (gp.py:362)."""
    X = np.asarray(X, dtype=np.float64); xt = np.atleast_2d(np.asarray(xt, dtype=np.float64))
    K = train_kernel(X, ell, sigma_f, sigma_n)
    K_inv = np.linalg.inv(K)
    a, b = xt / ell, X / ell
    d2 = ((a[:, None, :] - b[None, :, :]) ** 2).sum(axis=2)
    k_s = sigma_f * np.exp(-0.5 * d2)
    d2s = ((a[:, None, :] - a[None, :, :]) ** 2).sum(axis=2)
    k_ss = sigma_f * np.exp(-0.5 * d2s) + 1e-8 * np.eye(xt.shape[0])
    mu = k_s @ (K_inv @ y) + y_mean
    cov = k_ss - k_s @ K_inv @ k_s.T
    return mu, np.diag(cov).copy(), cov
