"""Small end-to-end exercise of every kernel (target for compute-sanitizer)."""
import math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import BatchSolver, default_opts, workload as wl
for N, B, M, variant in ((20, 37, 24, "4"), (40, 9, 0, "7"), (20, 5, 0, "1")):
    os.environ["ADMPC_QP_VARIANT"] = variant
    batch = wl.make_batch(B, N, seed=1, p=1.0, perturb=3.0)
    s = BatchSolver(B, default_opts(N))
    if M:
        s.set_gp(wl.make_gp(M=M, seed=2))
    L = 300
    sa = np.arange(L) * 0.5
    th = sa / 50.0
    traj = np.stack([np.full(L, 8.0), 50 * np.cos(th), 50 * np.sin(th), (th + math.pi / 2 + math.pi) % (2 * math.pi) - math.pi, sa, np.full(L, 0.02)], axis=1)
    s.set_track(traj, H=N, traj_dt=0.05, anchor=True)
    s.set_iterate(batch["x_init"], batch["u_init"])
    u, x, st = s.solve_batch(batch["x0"], batch["yref"], batch["p"][:, 0])
    s.closed_loop(2, use_track=True, log=True)
    s.get_lin(); s.get_lam(); s.get_t(); s.get_slacks(); s.get_waypoint_info()
    print("N", N, "variant", variant, "status ok", (st == 0).all(), flush=True)
    s.close()
print("DONE")
