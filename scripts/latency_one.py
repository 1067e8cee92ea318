"""Single-instance latency through the acados-shim symbols (cfg 1: nominal, N=20, B=1; development helper).
ADMPC_NO_GRAPH=1 keeps the eager stream path instead of the CUDA-graph replay."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import AcadosOcpSolverB200, default_opts, workload as wl
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
b1 = wl.make_batch(1, N, seed=20261, p=0.0)
cap = AcadosOcpSolverB200(default_opts(N))
for j in range(N):
    cap.set(j, "yref", b1["yref"][0][j * 9:(j + 1) * 9])
cap.set(N, "yref", b1["yref"][0][N * 9:])
for j in range(N + 1):
    cap.set(j, "x", b1["x_init"][0, j])
x0c = b1["x0"][0].copy()
ms, st = [], []
for it in range(205):
    cap.set(0, "lbx", x0c); cap.set(0, "ubx", x0c)
    t0 = time.perf_counter()
    st.append(cap.solve())
    ms.append((time.perf_counter() - t0) * 1e3)
    x0c = cap.get(1, "x")
ms = sorted(ms[5:])
print("single instance N=%d graph=%s: p50 %.4f ms  p99 %.4f ms  min %.4f  statuses ok %s  u0 %s" % (
    N, "off" if os.environ.get("ADMPC_NO_GRAPH") else "on", ms[len(ms) // 2], ms[int(0.99 * len(ms))], ms[0], all(s == 0 for s in st),
    np.array2string(cap.get(0, "u"), precision=12)))
