"""Reference-generator oracle pinned against the reference module's own outputs (tests/golden/refgen.npz)."""
import os

import numpy as np
import pytest

from oracle import refgen_oracle as ro


@pytest.mark.parametrize("H", [20, 40, 21])
def test_refgen_oracle_matches_reference_module(golden_dir, H):
    g = np.load(os.path.join(golden_dir, "refgen.npz"))
    traj, dt, pose = g["H%d_traj" % H], float(g["H%d_dt" % H]), g["H%d_pose" % H]
    for b in range(pose.shape[0]):
        w = ro.get_waypoints(traj, H, dt, *pose[b])
        for k in ("x_ref", "y_ref", "psi_ref", "v_ref", "cdist_ref", "curv_ref"):
            assert np.array_equal(w[k], g["H%d_%s" % (H, k)][b]), k        # same numpy calls: bit-exact
        for k in ("s0", "e_y0", "e_psi0"):
            assert w[k] == g["H%d_%s" % (H, k)][b]
        assert bool(w["stop"]) == bool(g["H%d_stop" % H][b])
