"""Kernel times of the Frenet variant (development helper). usage: frenet_time.py [B] [N] [M] [own] [spline]
own = 1: the variant's own constraint set (con_set = 1) ; spline = 1: kappa(s) spline inside the model (dense column of s)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import BatchSolver, default_opts, workload as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
M = int(sys.argv[3]) if len(sys.argv) > 3 else 200
OWN = int(sys.argv[4]) if len(sys.argv) > 4 else 0
SPL = int(sys.argv[5]) if len(sys.argv) > 5 else 0
b = wl.make_batch_frenet(B, N, seed=1, p=1.0)
if OWN:
    q = [0.0, 10.0, 10.0, 10.0, 10.0, 1.0, 0.1]
    opts = default_opts(N, model_variant=1, con_set=1, W=q + [10.0, 10.0], We=[0.01 * v for v in q], zl=[100.0, 100.0],
                        zu=[100.0, 100.0], lbu=[-10.0, -2.0], ubu=[5.0, 2.0], lbx=-0.52, ubx=0.52, lbx2=-2.0, ubx2=2.0)
else:
    opts = default_opts(N, model_variant=1)
s = BatchSolver(B, opts)
if SPL:
    from ad_mpc_b200 import kappa_pp_from_knots
    kn = np.linspace(-50.0, 450.0, 26)
    bb, cc = kappa_pp_from_knots(kn, 0.02 + 0.01 * np.sin(0.03 * kn))
    s.set_kappa_spline(np.tile(bb, (B, 1)), np.tile(cc, (B, 1, 1)))
if M:
    s.set_gp(wl.make_gp(M=M, seed=2))
s.set_profiling(True)
s.set_x0(b["x0"]); s.set_yref(b["yref"]); s.set_p(b["p"]); s.set_kappa(b["kappa"])
for r in range(4):
    s.set_iterate(b["x_init"], b["u_init"])
    s.solve(); s.wait()
st, qs, qi = s.get_status()
print("Frenet B=%d N=%d M=%d own=%d spline=%d variant=%s  prepare %.3f ms  qp %.3f ms  update %.3f ms  solve %.3f ms  -> %.3f Msolves/s (iters %.2f ok %d)" % (
    B, N, M, OWN, SPL, os.environ.get("ADMPC_QP_VARIANT", "0"), s.last_ms("prepare"), s.last_ms("qp"), s.last_ms("update"), s.last_ms("solve"), B / s.last_ms("solve") / 1e3, qi.mean(), (st == 0).all()))
s.close()
