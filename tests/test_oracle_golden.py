"""Pins the CPU oracle against the reference's own artefacts:

  * tests/golden/vde_vectors.npz   outputs of the reference's CasADi-generated sim_car_expl_vde_forw / _ode_fun
  * tests/golden/sim_car_iterate.npz  converged acados iterate (N=40, p=0): dynamics gap, KKT relations, multiplier
    conventions, and an RTI fixed-point test.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc

TS = 0.05


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_model_matches_reference_vde_vectors(golden_dir):
    """analytic f, Jx, Ju of the oracle == reference CasADi VDE (rows 0..5 of dSx; dSu pattern of the generated code)."""
    g = _load(golden_dir, "vde_vectors.npz")
    o = orc.default_opts()
    worst = 0.0
    for i in range(g["x"].shape[0]):
        f, Jx, Ju = orc.model_jac(o, g["x"][i], g["u"][i], g["p"][i])
        dSx = Jx @ g["Sx"][i]
        dSu = Jx @ g["Su"][i] + Ju
        scale = lambda a: np.maximum(1.0, np.abs(a))
        worst = max(worst, (np.abs(f - g["xdot"][i]) / scale(g["xdot"][i])).max())
        worst = max(worst, (np.abs(f - g["ode"][i]) / scale(g["ode"][i])).max())
        worst = max(worst, (np.abs(dSx[:6] - g["dSx"][i][:6]) / scale(g["dSx"][i][:6])).max())
        worst = max(worst, (np.abs(dSu[:6, 0] - g["dSu"][i][:6, 0]) / scale(g["dSu"][i][:6, 0])).max())
        worst = max(worst, (np.abs(dSu[:, 1] - g["dSu"][i][:, 1]) / scale(g["dSu"][i][:, 1])).max())
        # structural zeros of the generated code: row 6 of dSx, dSu[6,0]
        assert np.all(dSx[6] == 0.0) and dSu[6, 0] == 0.0
    assert worst < 1e-12, worst


def test_ref_backend_equals_analytic_backend():
    """RK4+sensitivities driven by the compiled reference VDE == driven by the analytic restatement."""
    if not orc.load_ref_model():
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    rng = np.random.default_rng(1)
    o = orc.default_opts()
    for p in (0.0, 1.0, 0.4):
        for _ in range(40):
            x = rng.normal(size=7) * np.array([10, 10, 1, 1, 0.5, 0.3, 0.2])
            x[3] = rng.uniform(1, 15)
            u = rng.normal(size=2)
            o.model_backend = 0
            xa, Aa, Ba, _ = orc.rk4_sens(o, x, u, p)
            o.model_backend = 1
            xr, Ar, Br, _ = orc.rk4_sens(o, x, u, p)
            for a, r in ((xa, xr), (Aa, Ar), (Ba, Br)):
                assert np.all(np.abs(a - r) <= 1e-12 * np.maximum(1.0, np.abs(r)))


def _golden_lin(g, o):
    N = g["u"].shape[0]
    A, B, gap = [], [], 0.0
    for k in range(N):
        xn, Ak, Bk, bad = orc.rk4_sens(o, g["x"][k], g["u"][k], 0.0)
        assert not bad
        gap = max(gap, np.abs(xn - g["x"][k + 1]).max())
        A.append(Ak)
        B.append(Bk)
    return np.array(A), np.array(B), gap


def _golden_lam(g):
    """[lbu0 lbu1 lbx | ubu0 ubu1 ubx | ls | us] for every stage; stage 0 has 7 state bounds (x0) instead of 1."""
    l0, t0 = g["lam0"], g["t0"]
    pick = lambda v: np.r_[v[0:2], 0.0, v[9:11], 0.0, v[18:22]]
    return np.vstack([pick(l0)[None], g["lam"]]), np.vstack([pick(t0)[None], g["t"]])


def test_golden_dynamics_gap(golden_dir):
    g = _load(golden_dir, "sim_car_iterate.npz")
    o = orc.default_opts(N=40)
    _, _, gap = _golden_lin(g, o)
    assert gap < 1e-12, gap          # 5.7e-14: pins f, RK4, dt=0.05, 1 step x 4 stages


def test_golden_kkt_relations(golden_dir):
    g = _load(golden_dir, "sim_car_iterate.npz")
    o = orc.default_opts(N=40)
    A, B, _ = _golden_lin(g, o)
    lam, t = _golden_lam(g)
    N = 40
    R = np.array([o.W[7], o.W[8]])
    for k in range(N):
        # u-stationarity: Ts R u + B^T pi - lam_lbu + lam_ubu = 0  (u_ref = 0)  -> pins B_k, Ts scaling, sign of pi
        s = TS * R * g["u"][k] + B[k].T @ g["pi"][k] - lam[k][0:2] + lam[k][3:5]
        assert np.abs(s).max() < 2e-10
        # slack stationarity: Ts z - lam_bound - lam_slack = 0
        assert np.abs(TS * 10.0 - lam[k][0:2] - lam[k][6:8]).max() < 1e-9
        assert np.abs(TS * 10.0 - lam[k][3:5] - lam[k][8:10]).max() < 1e-9
        # t definitions
        assert np.abs(t[k][0:2] - (g["u"][k] - np.array([-10, -3.0]) + g["sl"][k])).max() < 1e-12
        assert np.abs(t[k][3:5] - (np.array([5, 3.0]) - g["u"][k] + g["su"][k])).max() < 1e-12
        if k >= 1:
            assert abs(t[k][2] - (g["x"][k][6] + 0.52)) < 1e-12 and abs(t[k][5] - (0.52 - g["x"][k][6])) < 1e-12
    for k in range(1, N):
        # x-stationarity rows with zero weight (v_x, v_y, r, delta): A^T pi_k - pi_{k-1} -/+ lam_x = 0 -> pins A_k
        s = A[k].T @ g["pi"][k] - g["pi"][k - 1]
        s[6] += -lam[k][2] + lam[k][5]
        assert np.abs(s[3:]).max() < 1e-9
    # stage-0 multiplier of the eliminated x0 bound: lam_lbx0 - lam_ubx0 = q_0 + A_0^T pi_0 (here on rows 3..6, W=0)
    nu0 = g["lam0"][2:9] - g["lam0"][11:18]
    assert np.abs((A[0].T @ g["pi"][0])[3:] - nu0[3:]).max() < 1e-9


def _recover_yref(g, o, A, lam, We):
    """positions/heading reference from x-stationarity rows 0..2 (W>0); other rows copy the iterate (W=0)."""
    N = 40
    yref = np.zeros(N * 9 + 7)
    W = np.array(o.W[:7])
    for k in range(N):
        xr = g["x"][k].copy()
        if k >= 1:
            s = A[k].T @ g["pi"][k] - g["pi"][k - 1]
            xr[:3] = g["x"][k][:3] + s[:3] / (TS * W[:3])
        yref[k * 9:k * 9 + 7] = xr
    xr = g["x"][N].copy()
    xr[:3] = g["x"][N][:3] - g["pi"][N - 1][:3] / We[:3]
    yref[N * 9:] = xr
    return yref


def test_golden_rti_fixed_point(golden_dir):
    """One RTI step started AT the converged acados iterate must return (numerically) zero step and the golden duals."""
    g = _load(golden_dir, "sim_car_iterate.npz")
    We = np.array([10.0, 10.0, 100.0, 0, 0, 0, 0])   # the dump's W_e is not the in-tree one; any W_e with
    o = orc.default_opts(N=40, We=We)               # a consistent terminal reference gives the same fixed point
    A, B, _ = _golden_lin(g, o)
    lam, t = _golden_lam(g)
    yref = _recover_yref(g, o, A, lam, We)
    # recovered reference is a smooth path near the iterate
    ref_xy = yref[:40 * 9].reshape(40, 9)[:, :2]
    assert np.abs(ref_xy - g["x"][:40, :2]).max() < 2.0
    it = orc.make_iterate(o, g["x"], g["u"])
    st = orc.rti_step(o, it, g["x"][0], yref, np.zeros(40))
    assert st["status"] == 0 and st["qp_status"] == 0
    assert st["step_inf"] < 5e-7, st
    arr = orc.iterate_arrays(o, it)
    assert np.abs(arr["x"] - g["x"]).max() < 5e-7 and np.abs(arr["u"] - g["u"]).max() < 5e-7
    assert np.abs(arr["pi"] - g["pi"]).max() < 1e-5
    act = lam > 1e-3          # active multipliers (input bound at stage 0, slack multipliers 0.5 = Ts*z)
    assert np.abs(arr["lam"][act] - lam[act]).max() < 1e-5


def test_sqp_mode_oracle_properties(golden_dir):
    """Full SQP (nlp_solver_type "SQP", create_ros_ad_mpc.py:47-51): converges on the benchmark workload, a budget of k
    iterations equals k chained RTI steps bit for bit, and the acados golden iterate (a converged solution) passes the
    NLP residual check immediately (0 QPs)."""
    from ad_mpc_b200 import workload as wl
    N = 20
    b = wl.make_batch(12, N, seed=9, p=1.0, perturb=2.0)
    o = orc.default_opts(N)
    r = orc.sqp_batch(o, b["x0"], b["yref"], b["p"], b["x_init"], b["u_init"])
    assert (r["status"] == 0).all() and (r["sqp_iter"] >= 2).all() and (r["sqp_iter"] <= 12).all()
    assert (r["res"] < 1e-6).all()
    # one more RTI step from the converged point is (numerically) a zero step
    again = orc.rti_batch(o, b["x0"], b["yref"], b["p"], r["x"], r["u"])
    assert np.abs(again["u"] - r["u"]).max() < 1e-5 and np.abs(again["x"] - r["x"]).max() < 1e-5
    # budget of 2 iterations == 2 chained RTI steps, status MAXITER (2)
    r2 = orc.sqp_batch(o, b["x0"], b["yref"], b["p"], b["x_init"], b["u_init"], max_iter=2)
    s1 = orc.rti_batch(o, b["x0"], b["yref"], b["p"], b["x_init"], b["u_init"])
    s2 = orc.rti_batch(o, b["x0"], b["yref"], b["p"], s1["x"], s1["u"])
    assert (r2["status"] == 2).all() and (r2["sqp_iter"] == 2).all()
    assert np.array_equal(r2["x"], s2["x"]) and np.array_equal(r2["u"], s2["u"])
    # golden acados iterate: SQP from the converged point needs at most one QP
    g = _load(golden_dir, "sim_car_iterate.npz")
    We = np.array([10.0, 10.0, 100.0, 0, 0, 0, 0])
    o40 = orc.default_opts(N=40, We=We)
    A, B, _ = _golden_lin(g, o40)
    lam, t = _golden_lam(g)
    yref = _recover_yref(g, o40, A, lam, We)
    rg = orc.sqp_batch(o40, g["x"][0][None], yref[None], np.zeros((1, 40)), g["x"][None], g["u"][None], tol=(1e-5,) * 4)
    assert rg["status"][0] == 0 and rg["sqp_iter"][0] <= 1
    assert np.abs(rg["x"][0] - g["x"]).max() < 5e-6 and np.abs(rg["u"][0] - g["u"]).max() < 5e-6


def test_frenet_variant_oracle_consistency():
    """Frenet variant (SURVEY 8a A2', bytecode-only in the reference => no fixture): the analytic Jacobian matches central
    differences, and with zero curvature the variant IS the Cartesian model -- model, RK4 sensitivities and a whole RTI
    step reproduce the (reference-pinned) Cartesian oracle bit for bit."""
    from ad_mpc_b200 import workload as wl
    o, of = orc.default_opts(20), orc.default_opts(20, model_backend=2)
    rng = np.random.default_rng(0)
    for _ in range(40):
        x = np.array([rng.uniform(0, 50), rng.uniform(-1.5, 1.5), rng.uniform(-0.5, 0.5), rng.uniform(3, 14),
                      rng.uniform(-1, 1), rng.uniform(-0.6, 0.6), rng.uniform(-0.4, 0.4)])
        u, p, kap = rng.uniform(-2, 2, 2), rng.uniform(0, 1), rng.uniform(-0.1, 0.1)
        f, Jx, Ju = orc.model_jac(of, x, u, p, kappa=kap)
        h = 1e-6
        for j in range(7):
            xp, xm = x.copy(), x.copy()
            xp[j] += h; xm[j] -= h
            fd = (orc.model_jac(of, xp, u, p, kappa=kap)[0] - orc.model_jac(of, xm, u, p, kappa=kap)[0]) / (2 * h)
            assert np.abs(fd - Jx[:, j]).max() <= 1e-7 * max(1.0, np.abs(Jx[:, j]).max())
        f0, Jx0, Ju0 = orc.model_jac(of, x, u, p, kappa=0.0)
        fc, Jxc, Juc = orc.model_jac(o, x, u, p)
        assert np.array_equal(f0, fc) and np.array_equal(Jx0, Jxc) and np.array_equal(Ju0, Juc)
        # curvature enters rows 0 and 2 only
        assert np.array_equal(f[[1, 3, 4, 5, 6]], fc[[1, 3, 4, 5, 6]]) and f[0] != fc[0]
    b = wl.make_batch(8, 20, seed=3, p=1.0)
    r0 = orc.rti_batch(o, b["x0"], b["yref"], b["p"], b["x_init"], b["u_init"])
    r1 = orc.rti_batch(of, b["x0"], b["yref"], b["p"], b["x_init"], b["u_init"], kappa=np.zeros((8, 20)))
    assert np.array_equal(r0["u"], r1["u"]) and np.array_equal(r0["x"], r1["x"]) and np.array_equal(r0["qp_iter"], r1["qp_iter"])
    # curved path: the RTI step converges on the Frenet workload and tracks the path (e_y, e_psi shrink)
    bf = wl.make_batch_frenet(8, 20, seed=4, p=1.0, perturb=2.0)
    rf = orc.rti_batch(of, bf["x0"], bf["yref"], bf["p"], bf["x_init"], bf["u_init"], kappa=bf["kappa"])
    assert (rf["status"] == 0).all() and (rf["qp_status"] == 0).all()
    assert np.abs(rf["x"][:, -1, 1]).mean() < np.abs(bf["x0"][:, 1]).mean()


def test_frenet_spline_curvature_jacobian_matches_finite_differences():
    """Frenet backend with kappa(s) evaluated inside the model: the analytic Jacobian (including the d kappa / d s column)
    against central differences of f, and the RK4 sensitivities against differences of the RK4 map."""
    o = orc.default_opts(N=20)
    o.model_backend = 2
    kn = np.linspace(-5.0, 45.0, 11)
    from scipy.interpolate import CubicSpline
    cs = CubicSpline(kn, 0.03 * np.sin(kn / 7.0) + 0.01, bc_type="not-a-knot")
    breaks, coef = cs.x.copy(), np.ascontiguousarray(cs.c[::-1].T)
    orc.set_kappa_spline(breaks, coef)
    try:
        rng = np.random.default_rng(3)
        for _ in range(20):
            x = np.array([rng.uniform(0, 40), rng.normal() * 0.5, rng.normal() * 0.1, 8 + rng.normal(), rng.normal() * 0.3,
                          rng.normal() * 0.2, rng.normal() * 0.1])
            u = rng.normal(size=2) * np.array([1.0, 0.2])
            f, Jx, Ju = orc.model_jac(o, x, u, 1.0)
            assert abs(Jx[0, 0]) > 1e-6          # the column is live
            for j in range(7):
                e = np.zeros(7); e[j] = 1e-6
                fp = orc.model_jac(o, x + e, u, 1.0)[0]; fm = orc.model_jac(o, x - e, u, 1.0)[0]
                assert np.abs((fp - fm) / 2e-6 - Jx[:, j]).max() < 1e-6 * max(1.0, np.abs(Jx[:, j]).max())
            A = orc.rk4_sens(o, x, u, 1.0)[1]
            for j in range(7):
                e = np.zeros(7); e[j] = 1e-6
                d = (orc.rk4_sens(o, x + e, u, 1.0)[0] - orc.rk4_sens(o, x - e, u, 1.0)[0]) / 2e-6
                assert np.abs(d - A[:, j]).max() < 1e-6
    finally:
        orc.set_kappa_spline(None)


def test_frenet_constraint_set_matches_the_reference_iterate_dump(golden_dir):
    """ad_mpc/debug.json is an acados store_iterate dump of the Frenet variant (N = 40, a failed solve): it pins the STRUCTURE of
    that variant's inequality set -- 12 multipliers [lbu0 lbu1 lbx_ey lbx_delta | ub.. | ls0 ls1 | us0 us1] and 2 + 2 slacks per
    stage, 20 / 1 + 1 at stage 0 (initial state as bounds), none at the terminal node -- and, through the t-vector definitions and
    the slack stationarity lam_bound + lam_slack = Ts z, WHICH rows are soft (acceleration and steering angle) and the bounds
    (u0 in [-10, 5], u1 in [-2, 2], e_y in [-2, 2], delta in [-0.52, 0.52], zl = zu = 100).  The oracle's con_set = 1 must
    reproduce all of that from the dump's own x, u, lam, t, sl, su."""
    g = _load(golden_dir, "frenet_debug.npz")
    N = 40
    assert g["lam"].shape == (N - 1, 12) and g["sl"].shape == (N - 1, 2) and g["lam0"].shape == (20,) and g["sl0"].shape == (1,)
    o = orc.default_opts(N, con_set=1, model_backend=2, zl=[100.0, 100.0], zu=[100.0, 100.0], lbu=[-10.0, -2.0], ubu=[5.0, 2.0],
                         lbx=-0.52, ubx=0.52, lbx2=-2.0, ubx2=2.0)
    assert orc.con_rows(o) == 12
    lam, t = np.zeros((N, 12)), np.ones((N, 12))
    lam[1:], t[1:] = g["lam"], g["t"]
    # stage 0 of the dump: [lbu0 lbu1 lbx(7) | ubu0 ubu1 ubx(7) | ls0 | us0] ; this repo eliminates x0, the input rows remain
    for ours, theirs in ((0, 0), (1, 1), (4, 9), (5, 10), (8, 18), (10, 19)):
        lam[0, ours], t[0, ours] = g["lam0"][theirs], g["t0"][theirs]
    sl, su = np.zeros((N, 2)), np.zeros((N, 2))
    sl[1:], su[1:] = g["sl"], g["su"]
    sl[0, 0], su[0, 0] = g["sl0"][0], g["su0"][0]
    it = orc.make_iterate(o, g["x"], g["u"])
    orc.set_iterate_duals(o, it, lam=lam, t=t, sl=sl, su=su)
    e_t, e_s, e_c = orc.con_check(o, it)
    assert e_t < 1e-12, e_t                  # t = v - lo + sl, hi - v + su, sl, su with this row order and these bounds
    assert e_s < 1e-12, e_s                  # Ts z - lam_bound - lam_slack = 0: slack 0 <-> u0 rows, slack 1 <-> steering-angle rows
    assert e_c < 1e-8                        # the dump is complementary (lam t <= 7e-10)
    # the mapping is the only one that fits: with the soft state bound on e_y instead (SURVEY's first reading) stage 11 of
    # the dump (active lower steering bound, lam = 3.59, lam_ls1 = 1.41) violates the slack stationarity
    assert g["lam"][10, 3] > 3.0 and abs(g["lam"][10, 3] + g["lam"][10, 9] - 5.0) < 1e-12
    # and the hard rows carry multipliers with no slack partner (steering rate at its upper bound, stages 12..19)
    assert (g["lam"][11:19, 5] > 0.05).all()
    # Cartesian set on the same data must NOT fit (10 rows, different softness): guards against a vacuous check
    o0 = orc.default_opts(N, model_backend=2, zl=[100.0, 100.0], zu=[100.0, 100.0], lbu=[-10.0, -2.0], ubu=[5.0, 2.0])
    it0 = orc.make_iterate(o0, g["x"], g["u"])
    orc.set_iterate_duals(o0, it0, lam=lam[:, [0, 1, 3, 4, 5, 7, 8, 9, 10, 11]], t=t[:, [0, 1, 3, 4, 5, 7, 8, 9, 10, 11]], sl=sl, su=su)
    assert orc.con_check(o0, it0)[1] > 1.0
