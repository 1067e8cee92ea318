"""One batched RTI step (profiling target). usage: one.py B N M"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import BatchSolver, default_opts, workload as wl
B, N, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
batch = wl.make_batch(B, N, seed=1, p=1.0)
s = BatchSolver(B, default_opts(N))
if M:
    s.set_gp(wl.make_gp(M=M, seed=2))
s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"][:, 0])
for r in range(reps):
    s.set_iterate(batch["x_init"], batch["u_init"])
    s.solve(); s.wait()
st, qs, qi = s.get_status()
print("ok", (st == 0).all(), "iters", qi.mean(), "solve ms", s.last_ms("solve"))
s.close()
