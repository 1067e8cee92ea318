import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import PipelinedSolver, PinnedArray, default_opts, workload as wl
B, N = 16384, 20
batch = wl.make_batch(B, N, seed=20263, p=1.0)
model = wl.make_gp(M=200, seed=20263)
pin = {k: PinnedArray(v.shape) for k, v in (("x0", batch["x0"]), ("yref", batch["yref"]))}
pin["x0"].array[:] = batch["x0"]; pin["yref"].array[:] = batch["yref"]
pp = PinnedArray((B,)); pp.array[:] = 1.0
ou, ox, os_ = PinnedArray((B, N, 2)), PinnedArray((B, N + 1, 7)), PinnedArray((B,), dtype=np.int32)
for chunks in [int(c) for c in os.environ.get("CHUNKS", "1,2,4,6,8,10,12,16,24").split(",")]:
    ps = PipelinedSolver(B, default_opts(N), chunks=chunks)
    ps.set_gp(model)
    ts = []
    for it in range(8):
        ps.set_iterate(batch["x_init"], batch["u_init"]); ps.wait()
        t0 = time.perf_counter()
        ps.solve_batch(pin["x0"].array, pin["yref"].array, pp.array, ou.array, ox.array, os_.array)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("chunks", chunks, "e2e ms %.3f" % np.median(ts[3:]), "-> %.2f M/s" % (B / np.median(ts[3:]) / 1e3), flush=True)
    ps.close()
