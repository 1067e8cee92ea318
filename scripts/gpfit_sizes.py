"""GP fit evaluation times (NLL with / without alpha) at several training-set sizes (development helper)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import gpfit
rng = np.random.default_rng(0)
for M in (2000, 500, 200):
    X = rng.uniform(-1, 1, size=(M, 4)); y = np.sin(X[:, 0]) + 0.01 * rng.normal(size=M); y -= y.mean()
    for want in (True, False):
        for _ in range(3):
            nll, alpha, ms = gpfit.nll_alpha(X, y, np.ones(4) * 0.7, 0.5, 0.01, want_alpha=want)
        print("M=%d want_alpha=%s nll=%.6f device ms=%.3f" % (M, want, nll, ms))
