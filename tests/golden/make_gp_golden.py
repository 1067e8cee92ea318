"""Golden vectors for the GP path, produced by the reference's OWN numeric code (model_fitting/gp.py), imported in the
build container with its symbolic dependencies stubbed (casadi is only needed for the CasADi twins; utils.utils for
two unrelated helpers).  Pins: NLL values (gp.py:292-316), the fitted K^-1 y (gp.py:361-363) and the posterior mean of
`predict` (gp.py:426-430) and its posterior variance (return_cov, gp.py:425-441) on seeded data.

  python tests/golden/make_gp_golden.py   ->  tests/golden/gp_reference.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data_driven_mpc/ros_gp_mpc/src/model_fitting/gp.py"


def load_reference_gp():
    cs = types.ModuleType("casadi")

    class MX:                      # only used in isinstance checks on the numeric path
        pass

    class DM:                      # `self._K_cs = cs.DM(k)` mirrors; never read on the numeric path
        def __init__(self, *a, **k):
            pass

    cs.MX, cs.DM, cs.SX = MX, DM, MX
    sys.modules["casadi"] = cs
    utils = types.ModuleType("utils")
    uu = types.ModuleType("utils.utils")
    uu.safe_mknode_recursive = lambda *a, **k: None
    uu.make_bz_matrix = lambda *a, **k: None
    utils.utils = uu
    sys.modules["utils"] = utils
    sys.modules["utils.utils"] = uu
    if "tqdm" not in sys.modules:
        try:
            import tqdm  # noqa: F401
        except Exception:
            t = types.ModuleType("tqdm"); t.tqdm = lambda x, **k: x; sys.modules["tqdm"] = t
    spec = importlib.util.spec_from_file_location("gp_reference", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.CustomGPRegression.compute_gp_jac = lambda self: None     # CasADi-symbolic Jacobian: not part of the numeric path
    return mod


def main():
    gp = load_reference_gp()
    rng = np.random.default_rng(20264)
    out = {}
    for tag, M, dz in (("a", 80, 4), ("b", 60, 1)):
        lo = np.array([2.0, -1.0, -0.8, -0.52])[:dz]
        hi = np.array([14.0, 1.0, 0.8, 0.52])[:dz]
        X = rng.uniform(lo, hi, size=(M, dz))
        f = 0.3 * np.sin(X[:, min(1, dz - 1)]) + 0.02 * X[:, 0]
        y = f + 0.01 * rng.normal(size=M)
        y_mean = float(np.mean(y))
        kern = gp.CustomKernelFunctions("squared_exponential", params={"l": np.ones(dz) * 1.0, "sigma_f": 0.5})
        reg = gp.CustomGPRegression(list(range(dz)), [], 4, mean=np.zeros(dz), y_mean=y_mean, kernel=kern, sigma_n=0.01, n_restarts=1)
        # NLL at a few hyper-parameter vectors (log-space theta = [log l..., log sigma_f, log sigma_n])
        thetas = np.log(np.stack([np.r_[np.ones(dz) * s, 0.5, 0.01 * s] for s in (0.5, 1.0, 2.0)]))
        yc = y - y_mean
        nll_fun = reg._nll(X.copy(), yc.copy())
        out[tag + "_thetas"] = thetas
        out[tag + "_nll"] = np.array([nll_fun(t) for t in thetas])
        reg.fit(X.copy(), y.copy())                                   # L-BFGS-B on the NLL, then K, K^-1, K^-1 y
        xt = rng.uniform(lo, hi, size=(64, dz))
        mu = np.asarray(reg.predict(xt)).reshape(-1)
        mu_c, cov = reg.predict(xt, return_cov=True)                  # posterior variance (gp.py:425-441)
        assert np.allclose(np.asarray(mu_c).reshape(-1), mu)
        out[tag + "_var"] = np.asarray(cov).reshape(-1)
        out[tag + "_X"] = np.asarray(reg.x_train)
        out[tag + "_y"] = np.asarray(reg.y_train)                     # mean-subtracted by fit (gp.py:343)
        out[tag + "_y_mean"] = np.array(y_mean)
        out[tag + "_ell"] = np.asarray(reg.kernel.params["l"], dtype=np.float64).reshape(-1)
        out[tag + "_sigma_f"] = np.array(float(reg.kernel.params["sigma_f"]))
        out[tag + "_sigma_n"] = np.array(float(reg.sigma_n))
        out[tag + "_K_inv_y"] = np.asarray(reg.K_inv_y).reshape(-1)
        out[tag + "_xtest"] = xt
        out[tag + "_mu"] = mu
        print(tag, "M", M, "dz", dz, "ell", out[tag + "_ell"], "sigma_f", out[tag + "_sigma_f"], "sigma_n", out[tag + "_sigma_n"],
              "nll", out[tag + "_nll"], "mu[:3]", mu[:3])
    np.savez_compressed(os.path.join(HERE, "gp_reference.npz"), **out)


if __name__ == "__main__":
    main()
