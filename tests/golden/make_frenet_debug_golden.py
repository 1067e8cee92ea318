"""Golden fixture for the Frenet variant's constraint set, from the reference's own iterate dump (run in the build container only).

  python tests/golden/make_frenet_debug_golden.py

Input (read-only): /root/reference/data_driven_mpc/ros_gp_mpc/src/ad_mpc/debug.json -- an acados store_iterate dump of the
Frenet variant (fren_ad_3d_optimizer, N = 40; a failed solve, so it pins STRUCTURE and the multiplier / slack relations, not a
converged solution): per stage 12 multipliers [lbu0 lbu1 lbx_ey lbx_delta | ubu0 ubu1 ubx_ey ubx_delta | ls0 ls1 | us0 us1]
with 2 + 2 slacks, stage 0 with the 7 initial-state rows (20 multipliers, 1 + 1 slacks), no terminal rows.

Output: tests/golden/frenet_debug.npz   x[41,7] u[40,2] lam0[20] t0[20] sl0[1] su0[1] lam[39,12] t[39,12] sl[39,2] su[39,2]
"""
import json
import os

import numpy as np

REF = "/root/reference/data_driven_mpc/ros_gp_mpc/src/ad_mpc"
HERE = os.path.dirname(os.path.abspath(__file__))

d = json.load(open(os.path.join(REF, "debug.json")))
N = 40
assert d["lam_40"] == [] and d["u_40"] == [] and d["sl_40"] == []
out = dict(
    x=np.array([d["x_%d" % k] for k in range(N + 1)]), u=np.array([d["u_%d" % k] for k in range(N)]),
    lam0=np.array(d["lam_0"]), t0=np.array(d["t_0"]), sl0=np.array(d["sl_0"]), su0=np.array(d["su_0"]),
    lam=np.array([d["lam_%d" % k] for k in range(1, N)]), t=np.array([d["t_%d" % k] for k in range(1, N)]),
    sl=np.array([d["sl_%d" % k] for k in range(1, N)]), su=np.array([d["su_%d" % k] for k in range(1, N)]),
)
np.savez_compressed(os.path.join(HERE, "frenet_debug.npz"), **out)
print({k: v.shape for k, v in out.items()})
