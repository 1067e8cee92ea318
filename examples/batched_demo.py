"""Tour of the batched API on one GPU (run on a B200 box: `python examples/batched_demo.py`).

1. one RTI step for a batch of vehicles on a circular track (the reference's per-vehicle `run_optimization`, batched)
2. the same through the mirror of the reference's optimizer object
3. full SQP (the reference's point-reference mode)
4. GP-augmented model + on-device closed loop with reference generation, validity check and backup control
5. Frenet-frame model variant
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import AD3DOptimizerB200, BatchSolver, default_opts, workload as wl  # noqa: E402

B, N = 4096, 20
batch = wl.make_batch(B, N, seed=1, p=0.0)

# 1. plain batched RTI step
s = BatchSolver(B, default_opts(N))
s.set_iterate(batch["x_init"], batch["u_init"])
s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"])
s.solve()
status, qp_status, qp_iter = s.get_status()
print("1. RTI step: %d instances, %d ok, mean IPM iterations %.2f, first control %s" % (B, (status == 0).sum(), qp_iter.mean(), s.get_u()[0, 0]))

# 2. the reference's optimizer object, batched
opt = AD3DOptimizerB200(B=B, t_horizon=1.0, n_nodes=N)
opt.solver.set_iterate(batch["x_init"], batch["u_init"])
opt.set_reference_trajectory(batch["ref"], np.zeros((B, N + 1, 2)))
w_opt, x_opt, st = opt.run_optimization(initial_state=batch["x0"], return_x=True)
print("2. AD3DOptimizerB200.run_optimization: w_opt %s, x_opt %s, failures %d" % (w_opt.shape, x_opt.shape, (st != 0).sum()))
opt.close()

# 3. full SQP
s.set_iterate(batch["x_init"], batch["u_init"])
info = s.solve_sqp(max_iter=100)
print("3. full SQP: converged %d / %d in %d batch iterations (per instance %d..%d), worst residual %.1e" % (
    (info["status"] == 0).sum(), B, info["iterations_run"], info["sqp_iter"].min(), info["sqp_iter"].max(), info["res"].max()))

# 4. GP residual + closed loop on the device
s.set_gp(wl.make_gp(M=200, seed=2))
L = 800
th = np.arange(L) * (2 * math.pi / L)
track = np.stack([np.full(L, 8.0), 50 * np.cos(th), 50 * np.sin(th), (th + 1.5 * math.pi) % (2 * math.pi) - math.pi,
                  50.0 * th, np.full(L, 0.02)], axis=1)
s.set_track(track, H=N, traj_dt=0.05, anchor=True)
s.set_iterate(batch["x_init"], batch["u_init"]); s.set_x0(batch["x0"]); s.set_p(np.ones(B))
s.closed_loop(50)
out = s.get_loop_info()
print("4. closed loop, 50 control steps x %d vehicles: valid commands %.1f %%, mean speed %.2f m/s" % (
    B, 100.0 * out["valid"].mean(), out["x0"][:, 3].mean()))
s.close()

# 5. Frenet-frame variant
fb = wl.make_batch_frenet(B, N, seed=3, p=1.0)
f = BatchSolver(B, default_opts(N, model_variant=1))
f.set_iterate(fb["x_init"], fb["u_init"]); f.set_x0(fb["x0"]); f.set_yref(fb["yref"]); f.set_p(fb["p"]); f.set_kappa(fb["kappa"])
f.solve()
print("5. Frenet variant: %d ok, mean |e_y| at the end of the horizon %.3f m (start %.3f m)" % (
    (f.get_status()[0] == 0).sum(), np.abs(f.get_x()[:, -1, 1]).mean(), np.abs(fb["x0"][:, 1]).mean()))
f.close()

# 5b. the variant as the reference defines it: kappa(s) spline evaluated inside the model + its own constraint set
from ad_mpc_b200 import kappa_pp_from_knots
q = [0.0, 10.0, 10.0, 10.0, 10.0, 1.0, 0.1]
own = default_opts(N, model_variant=1, con_set=1, W=q + [10.0, 10.0], We=[0.01 * v for v in q], zl=[100.0, 100.0], zu=[100.0, 100.0],
                   lbu=[-10.0, -2.0], ubu=[5.0, 2.0], lbx=-0.52, ubx=0.52, lbx2=-2.0, ubx2=2.0)
s_knots = np.linspace(-50.0, 450.0, 26)
breaks, coef = kappa_pp_from_knots(s_knots, 0.02 + 0.01 * np.sin(0.03 * s_knots))     # what the reference hands to CasADi's bspline
f = BatchSolver(B, own)
f.set_kappa_spline(breaks, coef)                                                      # shared by all vehicles (or [B, ...] per vehicle)
f.set_iterate(fb["x_init"], fb["u_init"]); f.set_x0(fb["x0"]); f.set_yref(fb["yref"]); f.set_p(fb["p"])
f.solve()
print("5b. Frenet variant, spline curvature + own constraint set: %d ok, %d inequality rows per stage" % (
    (f.get_status()[0] == 0).sum(), f.get_lam().shape[-1]))
f.close()
