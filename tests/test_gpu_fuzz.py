"""Randomised option / scenario sweep: the CUDA path against the oracle on seeded random problem settings (weights,
bounds, slack penalties, horizon, time step, blend, perturbation size, IPM iteration budget), including settings that
drive the QP into active bounds, iteration limits and failures.  Status codes and iteration counts must be identical,
solutions within 1e-8 wherever the step was accepted."""
import numpy as np
import pytest

from ad_mpc_b200 import BatchSolver, default_opts, workload as wl
from oracle import oracle as orc
from util_parity import mirror_opts, mixed_err

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _random_setting(rng):
    N = int(rng.choice([5, 12, 20, 31, 32, 45, 70, 85]))
    q = list(rng.choice([0.0, 1.0, 10.0, 100.0], size=7))
    r = list(rng.choice([0.1, 1.0, 100.0], size=2))
    kw = dict(dt=float(rng.choice([0.02, 0.05, 0.1])), W=q + r, We=[float(rng.choice([1e-6, 1e-2, 1.0])) * v for v in q],
              zl=[float(rng.choice([1.0, 10.0, 100.0]))] * 2, zu=[float(rng.choice([1.0, 10.0, 100.0]))] * 2,
              Zl=[float(rng.choice([0.0, 0.5]))] * 2, Zu=[float(rng.choice([0.0, 2.0]))] * 2,
              lbu=[-float(rng.uniform(0.5, 10)), -float(rng.uniform(0.2, 3))], ubu=[float(rng.uniform(0.5, 5)), float(rng.uniform(0.2, 3))],
              lbx=-float(rng.uniform(0.05, 0.6)), ubx=float(rng.uniform(0.05, 0.6)),
              iter_max=int(rng.choice([2, 8, 50])))
    scen = dict(p=float(rng.choice([0.0, 0.4, 1.0])), perturb=float(rng.choice([0.5, 3.0, 12.0])))
    return N, kw, scen


@pytest.mark.parametrize("seed", list(range(16)))
def test_random_settings(seed):
    rng = np.random.default_rng(9000 + seed)
    N, kw, scen = _random_setting(rng)
    B = int(rng.choice([3, 17, 40]))
    batch = wl.make_batch(B, N, seed=100 + seed, dt=kw["dt"], **scen)
    batch["u_init"] = rng.normal(size=(B, N, 2)) * np.array([0.5, 0.1])          # inputs off zero: input bounds matter
    opts = default_opts(N, **kw)
    s = BatchSolver(B, opts)
    o = mirror_opts(opts)
    x, u = batch["x_init"], batch["u_init"]
    seen = set()
    for step in range(2):                                    # second step from the updated iterate
        s.set_iterate(x, u); s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"])
        s.solve()
        st, qs, qi = s.get_status()
        r = orc.rti_batch(o, batch["x0"], batch["yref"], batch["p"], x, u)
        assert np.array_equal(st, r["status"]), (seed, step, st, r["status"])
        assert np.array_equal(qs, r["qp_status"]), (seed, step)
        assert np.array_equal(qi, r["qp_iter"]), (seed, step, qi, r["qp_iter"])
        ok = r["status"] == 0
        gu, gx = s.get_u(), s.get_x()
        assert mixed_err(gu[ok], r["u"][ok]) <= TOL and mixed_err(gx[ok], r["x"][ok]) <= TOL, (seed, step)
        bad = ~ok                                            # rejected steps leave the iterate untouched on both sides
        assert np.array_equal(gu[bad], u[bad]) and np.array_equal(gx[bad], x[bad])
        seen.update(qs.tolist())
        x, u = r["x"], r["u"]
    s.close()
