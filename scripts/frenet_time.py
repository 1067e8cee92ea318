"""Kernel times of the Frenet variant's dense kernels (development helper). usage: frenet_time.py [B] [N] [M]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import BatchSolver, default_opts, workload as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
M = int(sys.argv[3]) if len(sys.argv) > 3 else 200
b = wl.make_batch_frenet(B, N, seed=1, p=1.0)
s = BatchSolver(B, default_opts(N, model_variant=1))
if M:
    s.set_gp(wl.make_gp(M=M, seed=2))
s.set_profiling(True)
s.set_x0(b["x0"]); s.set_yref(b["yref"]); s.set_p(b["p"]); s.set_kappa(b["kappa"])
for r in range(4):
    s.set_iterate(b["x_init"], b["u_init"])
    s.solve(); s.wait()
st, qs, qi = s.get_status()
print("Frenet B=%d N=%d M=%d  prepare %.3f ms  qp %.3f ms  update %.3f ms  solve %.3f ms  -> %.3f Msolves/s (iters %.2f ok %d)" % (
    B, N, M, s.last_ms("prepare"), s.last_ms("qp"), s.last_ms("update"), s.last_ms("solve"), B / s.last_ms("solve") / 1e3, qi.mean(), (st == 0).all()))
s.close()
