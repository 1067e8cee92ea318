"""QP / prepare kernel times of one configuration (development helper). usage: qp_time.py B N M [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import BatchSolver, default_opts, workload as wl
B, N, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
batch = wl.make_batch(B, N, seed=1, p=1.0)
s = BatchSolver(B, default_opts(N))
if M:
    s.set_gp(wl.make_gp(M=M, seed=2))
s.set_profiling(True)
s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"][:, 0])
tp, tq = [], []
for r in range(reps + 2):
    s.set_iterate(batch["x_init"], batch["u_init"])
    s.solve(); s.wait()
    if r >= 2:
        tp.append(s.last_ms("prepare")); tq.append(s.last_ms("qp"))
st, qs, qi = s.get_status()
print("%s B=%d N=%d M=%d  prepare %.3f ms  qp %.3f ms  (iters %.2f ok %d)" % (
    os.environ.get("TAG", ""), B, N, M, np.mean(tp), np.mean(tq), qi.mean(), (st == 0).all()), flush=True)
s.close()
