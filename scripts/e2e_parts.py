"""Where the end-to-end call spends its time beyond the device-resident step (development helper): the same pipelined call with
the uploads and / or the read-backs switched off (NULL pointers: the chunk handles keep their resident inputs / results)."""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ad_mpc_b200 import PipelinedSolver, PinnedArray, default_opts, workload as wl, _lib
B, N = 16384, 20
batch = wl.make_batch(B, N, seed=20263, p=1.0)
model = wl.make_gp(M=200, seed=20263)
pin = {k: PinnedArray(v.shape) for k, v in (("x0", batch["x0"]), ("yref", batch["yref"]))}
pin["x0"].array[:] = batch["x0"]; pin["yref"].array[:] = batch["yref"]
pp = PinnedArray((B,)); pp.array[:] = 1.0
ou, ox, os_ = PinnedArray((B, N, 2)), PinnedArray((B, N + 1, 7)), PinnedArray((B,), dtype=np.int32)
ps = PipelinedSolver(B, default_opts(N), chunks=8)
ps.set_gp(model)
L = _lib.load()
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
def run(up, down):
    ts = []
    for it in range(12):
        ps.set_iterate(batch["x_init"], batch["u_init"]); ps.wait()
        t0 = time.perf_counter()
        _lib.check(L.admpc_pipe_solve_host(ps.p, dp(pin["x0"].array) if up else None, dp(pin["yref"].array) if up else None,
                                           dp(pp.array) if up else None, dp(ou.array) if down else None, dp(ox.array) if down else None,
                                           ip(os_.array) if down else None), "solve_host")
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts[4:]))
run(True, True)
for up, down in ((True, True), (True, False), (False, True), (False, False)):
    print("upload %-5s read-back %-5s : %.3f ms per call" % (up, down, run(up, down)), flush=True)
ps.close()
