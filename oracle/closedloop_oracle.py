"""CPU ORACLE (test infrastructure only) for the after-solve logic and the closed loop.

Restates `is_valid_command` (reference: ad_mpc/ad_3d_optimizer.py:385-394), the backup-control rule (:469-476), the
safety counter of the node (nodes/gp_ad_mpc_node.py:62,206-216) and drives the RTI oracle + reference-generator oracle
in a loop with an RK4 plant of the nominal model.  Parity unpinned by reference fixtures (no recorded closed loop ships).
"""
import math

import numpy as np

from . import oracle as orc
from . import refgen_oracle as ro


def is_valid_command(x_opt, ref):                      # ad_3d_optimizer.py:385-394
    tmp = np.zeros(len(ref))
    for i in range(0, len(ref) - 1):
        tmp[i] = math.sqrt((ref[i, 0] - x_opt[i, 0]) ** 2 + (ref[i, 1] - x_opt[i, 1]) ** 2)
    return bool(np.mean(tmp) < 3.0 and np.cov(tmp) < 2 and np.max(tmp) < 4)


def closed_loop(o, traj, H, traj_dt, x0, p, x_init, u_init, steps, threshold=10, anchor=False, blend=None):
    """Returns (log[steps+1,B,7], info dict of the last step).  blend = (blend_min, blend_max): the kinematic/dynamic
    switch is recomputed from the measured v_x before every solve after the first (ad_3d_optimizer.py:443-450)."""
    B, N = x0.shape[0], o.N
    x0 = x0.copy()
    xit, uit = x_init.copy(), u_init.copy()
    prev_w = [None] * B
    cnt = np.zeros(B, dtype=np.int32)
    log = [x0.copy()]
    info = {}
    pfull = np.broadcast_to(np.asarray(p, dtype=np.float64).reshape(B, 1), (B, N)).copy()
    for _ in range(steps):
        yref, _ = ro.make_yref(traj, H, traj_dt, x0, N, anchor=anchor)
        r = orc.rti_batch(o, x0, yref, pfull, xit, uit)
        xit, uit = r["x"], r["u"]
        valid = np.zeros(B, dtype=np.int32)
        cmd_ok = np.zeros(B, dtype=np.int32)
        ua = np.zeros((B, 2))
        for b in range(B):
            ref = np.vstack([yref[b, :N * 9].reshape(N, 9)[:, :7], yref[b, N * 9:][None]])
            w = r["u"][b].reshape(-1)
            ok = is_valid_command(r["x"][b], ref)
            valid[b] = ok
            if ok:
                prev_w[b] = w.copy()
            elif prev_w[b] is not None:
                w = np.concatenate((prev_w[b][2:-1], prev_w[b][-3:-1]))          # (sic) ad_3d_optimizer.py:475
            cnt[b] = 0 if r["status"][b] > 0 else cnt[b] + 1
            cmd_ok[b] = int(cnt[b] >= threshold and ok)
            ua[b] = w[:2]
            u = np.array([w[0], min(max(w[1], o.lbu[1]), o.ubu[1])])
            x0[b] = orc.rk4_sens(o, x0[b], u, pfull[b, 0])[0]
            if blend is not None:
                pfull[b, :] = min(max((x0[b, 3] - blend[0]) / (blend[1] - blend[0]), 0.0), 1.0)
        log.append(x0.copy())
        info = dict(valid=valid, safe_count=cnt.copy(), cmd_ok=cmd_ok, u_apply=ua, status=r["status"])
    return np.array(log), info
