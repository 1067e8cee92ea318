"""CPU ORACLE (test infrastructure only) for the batched reference generator.

numpy restatement of `RefTrajectory.get_waypoints` (reference: ad_mpc/ref_traj.py:89-171, helpers :27-35) followed by
the node's reference assembly (nodes/gp_ad_mpc_node.py:180-187), the padding of `set_reference_trajectory`
(ad_mpc/ad_3d_optimizer.py:343-345) and the heading unwrap of `run_optimization` (:420-438).
Pinned against outputs of the reference module itself: tests/golden/refgen.npz (tests/golden/make_refgen_golden.py).
"""
import math

import numpy as np


def wrap_pi(a):
    """angle wrapped to [-pi, pi) with numpy's modulo convention (ref_traj.py:27-28)."""
    two_pi = 2.0 * np.pi
    return (a + np.pi) % two_pi - np.pi


def heading_relative_to(psi_ref, psi_vehicle):
    """reference headings made continuous around the vehicle heading (ref_traj.py:30-35)."""
    return psi_vehicle + np.unwrap(wrap_pi(psi_ref - psi_vehicle))


def track_tables(traj, H, dt):
    """Instance-independent part of get_waypoints (ref_traj.py:120-150): interpolation abscissae and tables."""
    vel = traj[:, 0]
    cd = traj[:, 4]
    while len(vel) < H + 1:
        vel = np.concatenate((vel, [0.01]))
    fit = [dt * vel[0]]
    for h in range(1, H):
        fit.append(fit[-1] + dt * vel[h])
    tab = {}
    for key, col in (("x", 1), ("y", 2), ("psi", 3), ("cdist", 4), ("curv", 5)):
        src = np.unwrap(traj[:, col]) if key == "psi" else traj[:, col]
        tab[key] = np.interp(fit, cd, src)
    v = np.diff(tab["cdist"]) / dt
    tab["v"] = np.insert(v, len(v), v[-1])
    tab["stop"] = bool(tab["cdist"][-1] == cd[-1])
    return tab


def track_tables_anchored(traj, H, dt, ci):
    """Extension (not in the reference): the same tables with arc length measured from waypoint `ci` of a shared global
    track and the speed profile taken from that waypoint on."""
    vel = traj[ci:, 0]
    cd = traj[:, 4]
    while len(vel) < H + 1:
        vel = np.concatenate((vel, [0.01]))
    fit = [0.0 + dt * vel[0]]
    for h in range(1, H):
        fit.append(fit[-1] + dt * vel[h])
    fit = traj[ci, 4] + np.array(fit)
    tab = {}
    for key, col in (("x", 1), ("y", 2), ("psi", 3), ("cdist", 4), ("curv", 5)):
        src = np.unwrap(traj[:, col]) if key == "psi" else traj[:, col]
        tab[key] = np.interp(fit, cd, src)
    v = np.diff(tab["cdist"]) / dt
    tab["v"] = np.insert(v, len(v), v[-1])
    tab["stop"] = bool(tab["cdist"][-1] == cd[-1])
    return tab


def get_waypoints(traj, H, dt, X0, Y0, psi0, anchor=False):
    """One pose -> dict like the reference's waypoint_dict (x_ref, y_ref, psi_ref, v_ref, s0, e_y0, e_psi0, stop)."""
    psi_init = wrap_pi(psi0)
    xy = traj[:, 1:3]
    ci = int(np.argmin(np.linalg.norm(xy - np.array([[X0, Y0]]), axis=1)))
    pw = traj[ci, 3]
    rot = np.array([[np.cos(pw), np.sin(pw)], [-np.sin(pw), np.cos(pw)]])
    ef = rot @ (np.array([X0, Y0]) - xy[ci])
    tab = track_tables_anchored(traj, H, dt, ci) if anchor else track_tables(traj, H, dt)
    psi = wrap_pi(heading_relative_to(tab["psi"], psi_init))
    out = dict(s0=traj[ci, 4], e_y0=ef[1], e_psi0=wrap_pi(psi_init - pw), stop=tab["stop"],
               cdist_ref=tab["cdist"], curv_ref=tab["curv"])
    out["x_ref"] = np.hstack([np.linspace(X0, tab["x"][1], 3), tab["x"][2:-1]])
    out["y_ref"] = np.hstack([np.linspace(Y0, tab["y"][1], 3), tab["y"][2:-1]])
    out["psi_ref"] = np.hstack([np.ones(3) * psi[0], psi[2:-1]])
    out["v_ref"] = np.hstack([np.ones(3) * tab["v"][2], tab["v"][2:-1]])
    return out


def make_yref(traj, H, dt, x0, N, anchor=False):
    """Batched: x0[B,7] -> yref[B, N*9+7] exactly as the node + optimizer assemble it."""
    B = x0.shape[0]
    yref = np.zeros((B, N * 9 + 7))
    info = np.zeros((B, 3))
    for b in range(B):
        w = get_waypoints(traj, H, dt, x0[b, 0], x0[b, 1], x0[b, 2], anchor=anchor)
        ref = np.zeros((H, 7))                                   # gp_ad_mpc_node.py:180-185
        ref[:, 0], ref[:, 1], ref[:, 2], ref[:, 3] = w["x_ref"], w["y_ref"], w["psi_ref"], w["v_ref"]
        while ref.shape[0] < N + 1:                              # ad_3d_optimizer.py:343-345
            ref = np.vstack((ref, ref[-1, :]))
        psi0 = x0[b, 2]
        for j in range(N + 1):                                   # ad_3d_optimizer.py:420-438
            r = ref[j].copy()
            if psi0 < 0:
                if psi0 + math.pi < r[2]:
                    r[2] -= 2 * math.pi
            elif psi0 > 0:
                if psi0 - math.pi > r[2]:
                    r[2] += 2 * math.pi
            if j < N:
                yref[b, j * 9:j * 9 + 7] = r
            else:
                yref[b, N * 9:] = r
        info[b] = (w["s0"], w["e_y0"], w["e_psi0"])
    return yref, info
