"""One GP fit evaluation (NLL + alpha) at M points -- ncu target. usage: gpfit_one.py [M]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import gpfit
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(0)
X = rng.uniform(-1, 1, size=(M, 4))
y = np.sin(X[:, 0]) + 0.01 * rng.normal(size=M)
y -= y.mean()
for _ in range(3):
    nll, alpha, ms = gpfit.nll_alpha(X, y, np.ones(4) * 0.7, 0.5, 0.01)
print("M=%d nll=%.6f device ms=%.3f" % (M, nll, ms))
