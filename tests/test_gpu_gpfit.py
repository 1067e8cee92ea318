"""GPU parity of the GP model update against the numpy oracle and the reference's own outputs."""
import os
import time

import numpy as np
import pytest

from ad_mpc_b200 import gpfit
from oracle import gpfit_oracle as gf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["a", "b"])
def test_nll_and_alpha_match_reference_outputs(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "gp_reference.npz"))
    X, y = g[tag + "_X"], g[tag + "_y"]
    dz = X.shape[1]
    for theta, ref in zip(g[tag + "_thetas"], g[tag + "_nll"]):
        p = np.exp(theta)
        nll, _, _ = gpfit.nll_alpha(X, y, p[:dz], p[dz], p[dz + 1], want_alpha=False)
        assert abs(nll - ref) <= 1e-9 * max(1.0, abs(ref))            # vs the reference's own _nll
    ell, sf, sn = g[tag + "_ell"], float(g[tag + "_sigma_f"]), float(g[tag + "_sigma_n"])
    nll, alpha, _ = gpfit.nll_alpha(X, y, ell, sf, sn)
    ref = g[tag + "_K_inv_y"]                                         # the reference's inv(K) @ y
    assert np.abs(alpha - ref).max() <= 1e-6 * np.abs(ref).max()
    K = gf.train_kernel(X, ell, sf, sn)
    assert np.abs(K @ alpha - y).max() <= 1e-9 * max(1.0, np.abs(y).max())


@pytest.mark.parametrize("M", [1, 31, 32, 33, 500, 2000])
def test_fit_sizes_against_oracle(M):
    rng = np.random.default_rng(M)
    dz = 4
    X = rng.uniform([2, -1, -0.8, -0.52], [14, 1, 0.8, 0.52], size=(M, dz))
    y = 0.3 * np.sin(X[:, 1]) + 0.01 * X[:, 0] * X[:, 3] + 0.01 * rng.normal(size=M)
    y = y - y.mean()
    ell, sf, sn = np.array([6.0, 1.0, 0.8, 0.5]), 0.5, 0.01
    nll, alpha, ms = gpfit.nll_alpha(X, y, ell, sf, sn)
    t0 = time.perf_counter()
    ref_nll = gf.nll(X, y, ell, sf, sn)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    assert abs(nll - ref_nll) <= 1e-8 * max(1.0, abs(ref_nll))
    K = gf.train_kernel(X, ell, sf, sn)
    assert np.abs(K @ alpha - y).max() <= 1e-8 * max(1.0, np.abs(y).max())
    if M == 2000:
        print("gp fit M=2000: device %.2f ms (numpy oracle %.1f ms), %.1f GFLOP/s on the factorisation" % (ms, cpu_ms, M ** 3 / 3 / ms / 1e6))


def test_fit_then_solve_round_trip():
    """fit on the GPU -> set_gp -> solve  ==  numpy-fitted model -> solve (the online model-update path)."""
    from ad_mpc_b200 import BatchSolver, default_opts, workload as wl
    rng = np.random.default_rng(9)
    M, dz = 96, 4
    X = rng.uniform(wl.GP_BOX_LO, wl.GP_BOX_HI, size=(M, dz))
    ys = [0.3 * np.sin(X[:, 1]) + 0.01 * rng.normal(size=M), 0.05 * X[:, 3] * X[:, 0] / 10 + 0.01 * rng.normal(size=M)]
    ell = 0.5 * (wl.GP_BOX_HI - wl.GP_BOX_LO)
    fits = [gpfit.fit(X, y, ell0=ell, sigma_f0=0.5, sigma_n0=0.01, optimise=False) for y in ys]
    model = gpfit.stack_models(fits)
    ref = dict(model, alpha=np.stack([gf.alpha(X, y - y.mean(), ell, 0.5, 0.01) for y in ys]))
    B, N = 32, 20
    batch = wl.make_batch(B, N, seed=5, p=1.0)
    out = []
    for m in (model, ref):
        s = BatchSolver(B, default_opts(N))
        s.set_gp(m)
        s.set_iterate(batch["x_init"], batch["u_init"])
        u, x, st = s.solve_batch(batch["x0"], batch["yref"], batch["p"][:, 0])
        out.append((u.copy(), st.copy()))
        s.close()
    assert np.array_equal(out[0][1], out[1][1]) and np.abs(out[0][0] - out[1][0]).max() < 1e-7


def test_hyperparameter_search_decreases_nll():
    rng = np.random.default_rng(11)
    M = 120
    X = rng.uniform(-2, 2, size=(M, 1))
    y = np.sin(2 * X[:, 0]) + 0.05 * rng.normal(size=M)
    m0 = gpfit.fit(X, y, sigma_f0=0.5, sigma_n0=0.01, optimise=False)
    m1 = gpfit.fit(X, y, sigma_f0=0.5, sigma_n0=0.01, optimise=True)
    assert m1["nll"] < m0["nll"] - 1.0
    yc = y - y.mean()
    ref = gf.nll(X, yc, m1["ell"], m1["sigma_f"], m1["sigma_n"])
    assert abs(ref - m1["nll"]) <= 1e-8 * max(1.0, abs(ref))
    assert np.all(m1["ell"] >= 1e-5 * (1 - 1e-9)) and np.all(m1["ell"] <= 10.0 * (1 + 1e-9)) and 1e-8 * (1 - 1e-9) <= m1["sigma_n"] <= 1.0 + 1e-9   # gp.py:336-338


@pytest.mark.parametrize("tag", ["a", "b"])
def test_posterior_mean_and_variance_match_reference_predict(golden_dir, tag):
    """admpc_gp_predict (Schur complement of the blocked Cholesky) against CustomGPRegression.predict(return_cov=True).
    The reference forms inv(K) (gp.py:362); the two routes agree to ~1e-12 absolute on variances of 1e-5."""
    g = np.load(os.path.join(golden_dir, "gp_reference.npz"))
    model = dict(X=g[tag + "_X"], y=g[tag + "_y"], ell=g[tag + "_ell"], sigma_f=float(g[tag + "_sigma_f"]),
                 sigma_n=float(g[tag + "_sigma_n"]), y_mean=float(g[tag + "_y_mean"]))
    mu, var, cov = gpfit.predict(model, g[tag + "_xtest"], full_cov=True)
    assert np.abs(mu - g[tag + "_mu"]).max() <= 1e-8 * max(1.0, np.abs(g[tag + "_mu"]).max())
    assert np.abs(var - g[tag + "_var"]).max() <= 1e-10
    _, _, cov_o = gf.predict(model["X"], model["y"], model["ell"], model["sigma_f"], model["sigma_n"], model["y_mean"], g[tag + "_xtest"])
    assert np.abs(cov - cov_o).max() <= 1e-10 and np.array_equal(cov, cov.T)
    assert np.array_equal(np.diag(cov), var)


@pytest.mark.parametrize("M,n", [(1, 1), (33, 70), (500, 257)])
def test_posterior_sizes_against_oracle(M, n):
    rng = np.random.default_rng(M + n)
    X = rng.uniform(-1, 1, size=(M, 3))
    y = np.sin(X[:, 0]) + 0.01 * rng.normal(size=M)
    y = y - y.mean()
    xt = rng.uniform(-1.5, 1.5, size=(n, 3))
    model = dict(X=X, y=y, ell=np.array([0.7, 1.1, 2.0]), sigma_f=0.8, sigma_n=0.05, y_mean=0.3)
    mu, var = gpfit.predict(model, xt, return_cov=True)
    mu_o, var_o, _ = gf.predict(X, y, model["ell"], 0.8, 0.05, 0.3, xt)
    assert np.abs(mu - mu_o).max() <= 1e-9 and np.abs(var - var_o).max() <= 1e-10
    assert (var > 0).all() and (var <= 0.8 + 1e-8 + 1e-12).all()
