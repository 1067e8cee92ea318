"""Golden vectors for the reference-trajectory generator, produced by IMPORTING the reference's own
`ad_mpc/ref_traj.py` in the build container (ROS-only imports are stubbed; the numerics are numpy/scipy).

  python tests/golden/make_refgen_golden.py   ->  tests/golden/refgen.npz

Reference: /root/reference/data_driven_mpc/ros_gp_mpc/src/ad_mpc/ref_traj.py:66-171 (set_traj, get_waypoints).
"""
import importlib.util
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data_driven_mpc/ros_gp_mpc/src/ad_mpc/ref_traj.py"


def load_reference_module():
    for name in ("rosbag", "rospy"):
        sys.modules.setdefault(name, types.ModuleType(name))     # imported at module top, unused by the numerics
    spec = importlib.util.spec_from_file_location("ref_traj_reference", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_track(L=600, ds=0.4, R=50.0, seed=3):
    """Waypoint list like the node receives: arc of the benchmark circle with a straight run-in, speeds 5..9 m/s."""
    rng = np.random.default_rng(seed)
    s = np.arange(L) * ds
    th = s / R - 2.9                                   # heading crosses +-pi inside the window (exercises the unwraps)
    x, y = R * np.cos(th), R * np.sin(th)
    psi = (th + math.pi / 2 + math.pi) % (2 * math.pi) - math.pi
    vel = 7.0 + 2.0 * np.sin(s / 30.0) + 0.05 * rng.normal(size=L)
    return x, y, psi, vel


def main():
    mod = load_reference_module()
    x, y, psi, vel = make_track()
    cases = {}
    for H, dt in ((20, 0.05), (40, 0.05), (21, 0.2)):
        rt = mod.RefTrajectory(traj_horizon=H, traj_dt=dt)
        rt.set_traj(x, y, psi, vel)
        traj = np.asarray(rt.trajectory)
        rng = np.random.default_rng(100 + H)
        B = 48
        idx = rng.integers(0, len(x) - 1, size=B)
        X0 = x[idx] + rng.normal(size=B) * 0.8
        Y0 = y[idx] + rng.normal(size=B) * 0.8
        P0 = psi[idx] + rng.normal(size=B) * 0.3
        P0[:6] += np.array([3.0, -3.0, 6.0, -6.0, 2 * math.pi, -2 * math.pi])     # unbounded inputs
        out = {k: [] for k in ("x_ref", "y_ref", "psi_ref", "v_ref", "s0", "e_y0", "e_psi0", "stop", "cdist_ref", "curv_ref")}
        for b in range(B):
            d = rt.get_waypoints(X0[b], Y0[b], P0[b])
            for k in out:
                out[k].append(np.asarray(d[k], dtype=np.float64))
        key = "H%d" % H
        cases[key + "_traj"] = traj
        cases[key + "_dt"] = np.array(dt)
        cases[key + "_pose"] = np.stack([X0, Y0, P0], axis=1)
        for k, v in out.items():
            cases[key + "_" + k] = np.stack(v)
    np.savez_compressed(os.path.join(HERE, "refgen.npz"), **cases)
    print({k: v.shape for k, v in cases.items() if k.startswith("H20")})


if __name__ == "__main__":
    main()
