"""Kernel-time sweep over batch sizes / variants (development helper, GPU box only)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import BatchSolver, default_opts, workload as wl

def run(B, N, gpM, variant, p=1.0, reps=5, **optkw):
    os.environ["ADMPC_QP_VARIANT"] = str(variant)
    batch = wl.make_batch(B, N, seed=1, p=p)
    s = BatchSolver(B, default_opts(N, **optkw))
    if gpM:
        s.set_gp(wl.make_gp(M=gpM, seed=2))
    s.set_profiling(True)
    s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"][:, 0])
    tp, tq, ts = [], [], []
    for r in range(reps + 2):
        s.set_iterate(batch["x_init"], batch["u_init"])
        s.solve(); s.wait()
        if r >= 2:
            tp.append(s.last_ms("prepare")); tq.append(s.last_ms("qp")); ts.append(s.last_ms("solve"))
    st, qs, qi = s.get_status()
    s.close()
    print("B=%6d N=%d M=%4d variant=%d  prepare %.3f ms  qp %.3f ms  solve %.3f ms  -> %.2f Msolves/s  (iters %.2f, ok %d)" % (
        B, N, gpM, variant, np.mean(tp), np.mean(tq), np.mean(ts), B / np.mean(ts) / 1e3, qi.mean(), (st == 0).all()), flush=True)

if __name__ == "__main__":
    variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4]
    for B in (1, 4096, 16384, 131072):
        for v in variants:
            run(B, 20, 0, v)
    run(16384, 20, 200, variants[-1])
    run(4096, 40, 2000, variants[-1], reps=2)
    print("# opt-in FP32-exponent GP (gp_precision = 1)")
    run(16384, 20, 200, variants[-1], gp_precision=1)
    run(4096, 40, 2000, variants[-1], reps=2, gp_precision=1)
