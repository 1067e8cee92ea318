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
