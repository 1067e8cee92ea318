"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerance (BASELINE.json north_star): 1e-8 relative in FP64, measured as |a-b| <= 1e-8*max(1,|b|); status codes and
QP iteration counts must be identical.
"""
import math
import os

import numpy as np
import pytest

from ad_mpc_b200 import BatchSolver, AcadosOcpSolverB200, default_opts, workload as wl
from oracle import oracle as orc
from util_parity import mirror_opts, mixed_err, oracle_batch

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _gpu_step(s, batch, gp_state=None):
    s.set_iterate(batch["x_init"], batch["u_init"])
    s.set_x0(batch["x0"])
    s.set_yref(batch["yref"])
    s.set_p(batch["p"])
    if gp_state is not None:
        s.set_gp_state(gp_state)
    s.solve()
    st, qs, qi = s.get_status()
    return dict(u=s.get_u(), x=s.get_x(), pi=s.get_pi(), status=st, qp_status=qs, qp_iter=qi)


def _compare(g, r, tol=TOL):
    assert np.array_equal(g["status"], r["status"])
    assert np.array_equal(g["qp_status"], r["qp_status"])
    assert np.array_equal(g["qp_iter"], r["qp_iter"]), (g["qp_iter"][:16], r["qp_iter"][:16])
    ok = r["status"] == 0
    assert mixed_err(g["u"][ok], r["u"][ok]) <= tol
    assert mixed_err(g["x"][ok], r["x"][ok]) <= tol
    assert mixed_err(g["pi"][ok], r["pi"][ok]) <= 1e-6      # duals: looser (scaled by 1/t near active bounds)


@pytest.mark.parametrize("p", [0.0, 1.0, 0.35])
@pytest.mark.parametrize("N", [20, 40])
def test_prepare_parity_nominal(p, N):
    B = 96
    batch = wl.make_batch(B, N, seed=11, p=p)
    rng = np.random.default_rng(5)
    batch["u_init"] = rng.normal(size=(B, N, 2)) * np.array([1.0, 0.2])      # exercise the u-dependent Jacobian
    opts = default_opts(N)
    s = BatchSolver(B, opts)
    s.set_iterate(batch["x_init"], batch["u_init"]); s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"])
    s.solve()
    lin = s.get_lin()
    o = mirror_opts(opts)
    for i in range(0, B, 7):
        it = orc.make_iterate(o, batch["x_init"][i], batch["u_init"][i])
        ref = orc.prepare(o, it, batch["yref"][i], batch["p"][i])
        for key in ("A", "B", "b", "q", "r"):
            assert mixed_err(lin[key][i], ref[key]) <= 1e-12, key
    s.close()


@pytest.mark.parametrize("p,trigger", [(1.0, 1), (0.0, 0), (0.6, 1)])
def test_prepare_parity_gp(p, trigger):
    B, N, M = 64, 20, 60
    batch = wl.make_batch(B, N, seed=12, p=p)
    rng = np.random.default_rng(6)
    batch["u_init"] = rng.normal(size=(B, N, 2)) * np.array([1.0, 0.2])
    model = wl.make_gp(M=M, seed=3)
    opts = default_opts(N)
    s = BatchSolver(B, opts)
    s.set_gp(model, stage0_trigger=trigger)
    gps = batch["x0"] + 0.01
    s.set_iterate(batch["x_init"], batch["u_init"]); s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"])
    s.set_gp_state(gps)
    s.solve()
    lin = s.get_lin()
    o = mirror_opts(opts)
    gp = orc.Gp(model)
    gp.apply(o, feat=model["feat"], rows=model["rows"], stage0_trigger=trigger)
    for i in range(0, B, 5):
        it = orc.make_iterate(o, batch["x_init"][i], batch["u_init"][i])
        ref = orc.prepare(o, it, batch["yref"][i], batch["p"][i], gp=gp, gp_state=gps[i])
        for key in ("A", "B", "b", "q", "r"):
            assert mixed_err(lin[key][i], ref[key]) <= 1e-10, key
    s.close()


@pytest.mark.parametrize("dz,n_out,M,feat,rows", [
    (1, 1, 37, (4,), (4,)),            # the reference's own 1-D models (README: --x 7 --y 7 style), odd M
    (2, 2, 1, (3, 8), (5, 3)),         # an input feature (u1), outputs on yaw rate and v_x, a single training point
    (3, 1, 130, (6, 7, 2), (5,)),      # steering angle, u0 and the heading as features
    (7, 2, 33, (2, 3, 4, 5, 6, 7, 8), (3, 4)),   # every admissible feature
])
def test_prepare_parity_gp_shapes(dz, n_out, M, feat, rows):
    """GP configurations off the benchmark's d_z = 4 fast path: generic feature loop, input features, 1..2 outputs."""
    B, N = 40, 20
    batch = wl.make_batch(B, N, seed=14, p=0.7)
    rng = np.random.default_rng(8)
    batch["u_init"] = rng.normal(size=(B, N, 2)) * np.array([1.0, 0.2])
    model = wl.make_gp(M=M, seed=4, n_out=n_out, dz=min(dz, 4))
    if dz > 4:                                   # widen the synthetic model to dz features
        model["X"] = np.concatenate([model["X"], rng.uniform(-1, 1, size=(n_out, M, dz - 4))], axis=2)
        model["ell"] = np.concatenate([model["ell"], rng.uniform(1.0, 3.0, size=(n_out, dz - 4))], axis=1)
    model["feat"], model["rows"] = feat, rows
    opts = default_opts(N)
    s = BatchSolver(B, opts)
    s.set_gp(model, stage0_trigger=1)
    gps = batch["x0"] + 0.02
    s.set_iterate(batch["x_init"], batch["u_init"]); s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"])
    s.set_gp_state(gps)
    s.solve()
    lin = s.get_lin()
    o = mirror_opts(opts)
    gp = orc.Gp(model)
    gp.apply(o, feat=feat, rows=rows, stage0_trigger=1)
    for i in range(0, B, 7):
        it = orc.make_iterate(o, batch["x_init"][i], batch["u_init"][i])
        ref = orc.prepare(o, it, batch["yref"][i], batch["p"][i], gp=gp, gp_state=gps[i])
        for key in ("A", "B", "b", "q", "r"):
            assert mixed_err(lin[key][i], ref[key]) <= 1e-10, key
    # and the whole step
    g = dict(u=s.get_u(), x=s.get_x())
    r = orc.rti_batch(o, batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], gp=gp, gp_state=gps)
    assert mixed_err(g["u"], r["u"]) <= TOL and mixed_err(g["x"], r["x"]) <= TOL
    s.close()


@pytest.mark.parametrize("p", [0.0, 1.0])
@pytest.mark.parametrize("B,N", [(1, 20), (33, 20), (512, 20), (100, 40)])
def test_rti_step_parity_nominal(p, B, N):
    batch = wl.make_batch(B, N, seed=20262, p=p)
    opts = default_opts(N)
    s = BatchSolver(B, opts)
    g = _gpu_step(s, batch)
    r = oracle_batch(mirror_opts(opts), batch)
    assert (r["status"] == 0).all()
    _compare(g, r)
    s.close()


def test_rti_step_parity_active_bounds():
    """large initial-state errors: acceleration / steering-rate soft bounds and the steering bound become active."""
    B, N = 256, 20
    batch = wl.make_batch(B, N, seed=99, p=0.0, perturb=6.0)
    batch["x0"][:, 6] = np.clip(batch["x0"][:, 6] * 4, -0.6, 0.6)        # some start outside the steering bound
    batch["x_init"][:, :, 6] = np.clip(batch["x0"][:, None, 6], -0.5, 0.5)
    opts = default_opts(N)
    s = BatchSolver(B, opts)
    g = _gpu_step(s, batch)
    r = oracle_batch(mirror_opts(opts), batch)
    lam_active = (np.abs(g["u"][..., 0] - 5.0) < 1e-6).any() or (np.abs(g["u"][..., 0] + 10.0) < 1e-6).any()
    assert lam_active, "test does not exercise an active input bound"
    _compare(g, r)
    s.close()


def test_rti_step_parity_gp():
    B, N, M = 128, 20, 200
    batch = wl.make_batch(B, N, seed=20263, p=1.0)
    model = wl.make_gp(M=M, seed=20263)
    opts = default_opts(N)
    s = BatchSolver(B, opts)
    s.set_gp(model)
    g = _gpu_step(s, batch)
    o = mirror_opts(opts)
    gp = orc.Gp(model)
    gp.apply(o, feat=model["feat"], rows=model["rows"])
    r = oracle_batch(o, batch, gp=gp)
    _compare(g, r)
    s.close()


def test_closed_loop_warm_start_parity():
    """5 consecutive RTI steps (iterate carried, unshifted, like the reference) stay in parity."""
    B, N = 64, 20
    batch = wl.make_batch(B, N, seed=7, p=1.0)
    opts = default_opts(N)
    s = BatchSolver(B, opts)
    o = mirror_opts(opts)
    xo, uo = batch["x_init"].copy(), batch["u_init"].copy()
    s.set_iterate(xo, uo)
    x0 = batch["x0"].copy()
    for step in range(5):
        u, x, st = s.solve_batch(x0, batch["yref"], batch["p"][:, 0])
        r = orc.rti_batch(o, x0, batch["yref"], batch["p"], xo, uo)
        assert np.array_equal(st, r["status"])
        assert mixed_err(u, r["u"]) <= TOL and mixed_err(x, r["x"]) <= TOL
        xo, uo = r["x"], r["u"]
        x0 = r["x"][:, 1, :].copy()          # plant = model prediction
    s.close()


def test_golden_fixed_point_on_gpu(golden_dir):
    """The converged acados iterate (sim_car_iterate.json) is a fixed point of the CUDA RTI step as well."""
    from test_oracle_golden import _golden_lam, _golden_lin, _recover_yref
    g = np.load(os.path.join(golden_dir, "sim_car_iterate.npz"))
    We = np.array([10.0, 10.0, 100.0, 0, 0, 0, 0])
    opts = default_opts(40, We=We)
    o = mirror_opts(opts)
    A, B, _ = _golden_lin(g, o)
    lam, t = _golden_lam(g)
    yref = _recover_yref(g, o, A, lam, We)
    s = BatchSolver(1, opts)
    s.set_iterate(g["x"][None], g["u"][None])
    u, x, st = s.solve_batch(g["x"][0][None], yref[None], np.zeros(1))
    assert st[0] == 0
    assert np.abs(x[0] - g["x"]).max() < 5e-7 and np.abs(u[0] - g["u"]).max() < 5e-7
    assert np.abs(s.get_pi()[0] - g["pi"]).max() < 1e-5
    lam_g = s.get_lam()[0]
    act = lam > 1e-3
    assert np.abs(lam_g[act] - lam[act]).max() < 1e-5
    s.close()


def test_load_iterate_restores_duals_and_sqp_sees_convergence(golden_dir, tmp_path):
    """acados' load_iterate restores x, u, pi, lam, t, sl, su (ad_3d_optimizer.py:454).  The reference's own converged
    iterate (sim_car_iterate.json, N = 40), written back in the acados JSON layout and loaded through the shim, must
    test as converged in SQP mode without a single QP (it would not if the multipliers were dropped)."""
    import json
    from test_oracle_golden import _golden_lam, _golden_lin, _recover_yref
    g = np.load(os.path.join(golden_dir, "sim_car_iterate.npz"))
    N = 40
    We = np.array([10.0, 10.0, 100.0, 0, 0, 0, 0])
    opts = default_opts(N, We=We)
    o = mirror_opts(opts)
    A, B, _ = _golden_lin(g, o)
    lam, t = _golden_lam(g)
    yref = _recover_yref(g, o, A, lam, We)
    d = {}
    for k in range(N + 1):
        d["x_%d" % k] = g["x"][k].tolist()
        d["z_%d" % k] = []
        if k < N:
            d["u_%d" % k] = g["u"][k].tolist(); d["pi_%d" % k] = g["pi"][k].tolist()
            d["lam_%d" % k] = (g["lam0"] if k == 0 else g["lam"][k - 1]).tolist()
            d["t_%d" % k] = (g["t0"] if k == 0 else g["t"][k - 1]).tolist()
            d["sl_%d" % k] = g["sl"][k].tolist(); d["su_%d" % k] = g["su"][k].tolist()
        else:
            for f in ("u", "lam", "t", "sl", "su"):
                d["%s_%d" % (f, k)] = []
    fn = str(tmp_path / "iterate.json")
    with open(fn, "w") as fh:
        json.dump(d, fh)

    def run(with_duals):
        cap = AcadosOcpSolverB200(opts, nlp_solver_type="SQP")
        for j in range(N):
            cap.set(j, "yref", yref[j * 9:(j + 1) * 9]); cap.set(j, "p", np.zeros(1))
        cap.set(N, "yref", yref[N * 9:])
        cap.set(0, "lbx", g["x"][0]); cap.set(0, "ubx", g["x"][0])
        if with_duals:
            cap.load_iterate(fn)
        else:
            for k in range(N + 1):
                cap.set(k, "x", g["x"][k])
                if k < N:
                    cap.set(k, "u", g["u"][k])
        st = cap.solve()
        return st, cap.get_stats("sqp_iter"), cap
    st, it, cap = run(True)
    assert st == 0 and it == 0, (st, it)
    # what was loaded comes back through the getters (stage-0 acados layout included)
    assert np.abs(cap.get(3, "pi") - g["pi"][3]).max() == 0 and np.abs(cap.get(5, "lam") - g["lam"][4]).max() == 0
    assert np.abs(cap.get(0, "lam")[[0, 1, 9, 10, 18, 19, 20, 21]] - g["lam0"][[0, 1, 9, 10, 18, 19, 20, 21]]).max() == 0
    st2, it2, _ = run(False)
    assert st2 == 0 and it2 >= 1            # without the multipliers the same point does not pass the KKT test


def test_cond_N_update_is_accepted():
    """acados_solver_sim_car.h:137 is exported; there is nothing to re-condense here, so it is a no-op returning 0."""
    from ad_mpc_b200 import _lib
    cap = AcadosOcpSolverB200(default_opts(20))
    assert _lib.load().sim_car_acados_update_qp_solver_cond_N(cap.c, 5) == 0


def test_nan_linearisation_status():
    """NaN in the iterate -> ACADOS_FAILURE (1) for that instance only; others unaffected."""
    B, N = 40, 20
    batch = wl.make_batch(B, N, seed=3)
    batch["x_init"][5, 3, 3] = np.nan
    opts = default_opts(N)
    s = BatchSolver(B, opts)
    g = _gpu_step(s, batch)
    r = oracle_batch(mirror_opts(opts), batch)
    assert g["status"][5] == 1 and r["status"][5] == 1
    assert np.array_equal(g["status"], r["status"])
    s.close()


def test_iter_max_status():
    """QP iteration limit: qp_status 2 (ACADOS_MAXITER), RTI status 0 (tolerated), same on both sides."""
    B, N = 32, 20
    batch = wl.make_batch(B, N, seed=4)
    opts = default_opts(N, iter_max=2)
    s = BatchSolver(B, opts)
    g = _gpu_step(s, batch)
    r = oracle_batch(mirror_opts(opts), batch)
    assert (g["qp_status"] == 2).all() and (g["qp_iter"] == 2).all()
    _compare(g, r, tol=1e-7)
    s.close()


def test_acados_shim_single_instance():
    """The acados-shim symbols reproduce AD3DOptimizer.run_optimization's call sequence (ad_3d_optimizer.py:420-465)."""
    N = 20
    batch = wl.make_batch(1, N, seed=8, p=0.0)
    opts = default_opts(N)
    solver = AcadosOcpSolverB200(opts)
    yref = batch["yref"][0]
    for j in range(N):
        solver.set(j, "yref", yref[j * 9:(j + 1) * 9])
    solver.set(N, "yref", yref[N * 9:])
    solver.set(0, "lbx", batch["x0"][0])
    solver.set(0, "ubx", batch["x0"][0])
    for j in range(N + 1):
        solver.set(j, "p", np.array([0.0]))
    for j in range(N + 1):
        solver.set(j, "x", batch["x_init"][0, j])
    status = solver.solve()
    assert status == 0
    u = np.array([solver.get(i, "u") for i in range(N)])
    x = np.array([solver.get(i, "x") for i in range(N + 1)])
    r = oracle_batch(mirror_opts(opts), batch)
    assert mixed_err(u, r["u"][0]) <= TOL and mixed_err(x, r["x"][0]) <= TOL
    assert solver.get_stats("sqp_iter") == 1 and solver.get_stats("qp_iter") == r["qp_iter"][0]
    assert solver.get(0, "lam").shape == (22,) and solver.get(1, "lam").shape == (10,)
    assert solver.get_stats("kkt_norm_inf") < 1e-8
    d = solver.store_iterate()
    assert len(d) == 8 * (N + 1) - 1
    # second solve from the carried iterate changes less than the first
    x_prev = x.copy()
    solver.solve()
    x2 = np.array([solver.get(i, "x") for i in range(N + 1)])
    assert np.abs(x2 - x_prev).max() < 0.5


def test_scale_invariance_full_size():
    """BASELINE cfg-3 size (B=16384, GP M=200): size-independent properties instead of a 16k-instance oracle run --
    (i) duplicated instances give bit-identical results wherever they sit in the batch, (ii) a 256-instance sample
    matches the oracle, (iii) every status is 0 and every QP converged."""
    B, N = 16384, 20
    base = wl.make_batch(256, N, seed=20263, p=1.0)
    reps = B // 256
    batch = {k: np.concatenate([v] * reps, axis=0) for k, v in base.items()}
    model = wl.make_gp(M=200, seed=20263)
    opts = default_opts(N)
    s = BatchSolver(B, opts)
    s.set_gp(model)
    g = _gpu_step(s, batch)
    assert (g["status"] == 0).all() and (g["qp_status"] == 0).all()
    u = g["u"].reshape(reps, 256, N, 2)
    assert np.array_equal(u[0], u[-1]) and np.array_equal(u[0], u[reps // 2])
    o = mirror_opts(opts)
    gp = orc.Gp(model)
    gp.apply(o, feat=model["feat"], rows=model["rows"])
    r = oracle_batch(o, base, gp=gp)
    assert mixed_err(u[0], r["u"]) <= TOL
    assert np.array_equal(g["qp_iter"][:256], r["qp_iter"])
    s.close()


def test_long_horizon_fallback_variants():
    """Horizon dispatch of the feedback kernel: N <= 31 the tensor-core kernel with one warp per instance (qp_mma<1>), 32..63
    with two (qp_mma<2>), 64..127 with four (qp_mma<4>), above that one thread per instance -- same parity bar on both
    sides of every edge."""
    launches = {}
    for N, B in ((31, 16), (32, 16), (40, 24), (63, 12), (64, 12), (90, 12), (127, 6), (128, 6)):
        batch = wl.make_batch(B, N, seed=31, p=1.0, perturb=3.0 if N in (32, 63, 127) else 1.0)
        opts = default_opts(N)
        s = BatchSolver(B, opts)
        n0 = s.kernel_launches()
        g = _gpu_step(s, batch)
        launches[N] = s.kernel_launches() - n0
        r = oracle_batch(mirror_opts(opts), batch)
        _compare(g, r)
        s.close()
    # the update is fused into the feedback kernel up to N = 127; the thread-per-instance kernel beyond needs one more launch
    assert launches[31] == launches[64] == launches[127] and launches[128] == launches[127] + 1


def test_multi_warp_kernels_match_the_other_implementations(monkeypatch):
    """N = 40 with active bounds: the two-warps-per-instance tensor-core kernel (default, 7) against the register-resident
    two-warp kernel (4) and the thread-per-instance kernel (1); N = 90: four warps per instance (7) against (1)."""
    for N, B, others in ((40, 96, (1, 4)), (90, 24, (1,))):
        batch = wl.make_batch(B, N, seed=78, p=0.5, perturb=5.0)
        out = {}
        for v in others + (7,):
            monkeypatch.setenv("ADMPC_QP_VARIANT", str(v))
            s = BatchSolver(B, default_opts(N))
            out[v] = _gpu_step(s, batch)
            s.close()
        for v in others:
            assert np.array_equal(out[v]["qp_iter"], out[7]["qp_iter"]) and np.array_equal(out[v]["status"], out[7]["status"])
            assert mixed_err(out[v]["u"], out[7]["u"]) <= TOL and mixed_err(out[v]["x"], out[7]["x"]) <= TOL
        r = oracle_batch(mirror_opts(default_opts(N)), batch)
        _compare(out[7], r)


def test_qp_variants_agree(monkeypatch):
    """The three QP kernels are independent implementations of the same algorithm: identical statuses / iteration
    counts and 1e-8 agreement on a batch with active bounds (7 = tensor-core sweeps, the default)."""
    B, N = 128, 20
    batch = wl.make_batch(B, N, seed=77, p=0.5, perturb=5.0)
    out = {}
    for v in (1, 4, 7):
        monkeypatch.setenv("ADMPC_QP_VARIANT", str(v))
        s = BatchSolver(B, default_opts(N))
        out[v] = _gpu_step(s, batch)
        s.close()
    for v in (1, 4):
        assert np.array_equal(out[v]["qp_iter"], out[7]["qp_iter"]) and np.array_equal(out[v]["status"], out[7]["status"]), v
        assert mixed_err(out[v]["u"], out[7]["u"]) <= TOL and mixed_err(out[v]["x"], out[7]["x"]) <= TOL, v


def test_non_default_weights_and_quadratic_slack_penalty():
    """Python-default weights of AD3DOptimizer (ad_3d_optimizer.py:45-47: q = [10,10,50,0,0,0,1], r = [1,100]),
    W_e = 1e-6 q (:151), plus a quadratic slack term (Zl, Zu > 0) and tighter input bounds."""
    B, N = 96, 20
    batch = wl.make_batch(B, N, seed=41, p=0.0, perturb=4.0)
    q = [10, 10, 50, 0.0, 0.0, 0.0, 1]
    opts = default_opts(N, W=q + [1.0, 100.0], We=[1e-6 * v for v in q], Zl=[0.5, 2.0], Zu=[0.5, 2.0],
                        lbu=[-3.0, -1.0], ubu=[2.0, 1.0])
    s = BatchSolver(B, opts)
    g = _gpu_step(s, batch)
    r = oracle_batch(mirror_opts(opts), batch)
    assert (np.abs(g["u"][..., 0]) > 1.99).any()          # the tighter acceleration bound is exercised
    _compare(g, r)
    s.close()


def test_error_paths():
    from ad_mpc_b200 import _lib
    import ctypes as C
    L = _lib.load()
    s = BatchSolver(8, default_opts(20))
    model = wl.make_gp(M=8, seed=1)
    bad = dict(model, feat=(0, 4, 5, 6))                   # p_x as a GP feature would break the A-structure
    with pytest.raises(_lib.AdmpcError):
        s.set_gp(bad)
    with pytest.raises(_lib.AdmpcError):
        s.make_yref()                                      # no track set
    assert L.admpc_batch_set_x0(s.h, None) == -1
    s.close()
    cap = AcadosOcpSolverB200(default_opts(20))
    with pytest.raises(_lib.AdmpcError):
        cap.set(3, "yref", np.zeros(5))                    # wrong length
    with pytest.raises(_lib.AdmpcError):
        cap.set(2, "lbx", np.zeros(1))                     # per-stage bound changes are not supported
    with pytest.raises(_lib.AdmpcError):
        cap.set(0, "nonsense", np.zeros(1))
    o = default_opts(20)
    h = C.c_void_p()
    assert L.admpc_batch_create(C.byref(o), 0, 0, C.byref(h)) == -1       # B must be positive
    o.N = 1000
    assert L.admpc_batch_create(C.byref(o), 4, 0, C.byref(h)) == -1       # N > ADMPC_NMAX


def test_c_example_runs():
    """The plain-C example (reference main_sim_car.c sequence) links against libadmpc_b200.so and converges."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "main_sim_car_b200")
    if not os.path.exists(exe):
        subprocess.run(["/usr/bin/gcc", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "main_sim_car_b200.c"),
                        "-o", exe, "-L", os.path.join(root, "ad_mpc_b200"), "-ladmpc_b200",
                        "-Wl,-rpath," + os.path.join(root, "ad_mpc_b200"), "-lm"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "EXAMPLE_OK" in out.stdout, out.stdout[-800:] + out.stderr[-400:]


def _gpu_sqp(s, batch, **kw):
    s.set_iterate(batch["x_init"], batch["u_init"])
    s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"])
    info = s.solve_sqp(**kw)
    info.update(u=s.get_u(), x=s.get_x())
    return info


@pytest.mark.parametrize("B,N,p", [(200, 20, 1.0), (64, 40, 0.0)])
def test_sqp_mode_parity(B, N, p):
    """Full SQP (point-reference mode of the reference, create_ros_ad_mpc.py:47-51): identical acados statuses and SQP
    iteration counts, 1e-8 agreement of the converged trajectories; instances finish at different iterations."""
    batch = wl.make_batch(B, N, seed=90 + N, p=p, perturb=3.0)
    opts = default_opts(N)
    s = BatchSolver(B, opts)
    g = _gpu_sqp(s, batch)
    r = orc.sqp_batch(mirror_opts(opts), batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"])
    assert np.array_equal(g["status"], r["status"]) and (r["status"] == 0).all()
    assert np.array_equal(g["sqp_iter"], r["sqp_iter"]), (g["sqp_iter"][:16], r["sqp_iter"][:16])
    assert len(set(r["sqp_iter"].tolist())) > 1 and g["iterations_run"] == r["sqp_iter"].max()
    assert mixed_err(g["u"], r["u"]) <= TOL and mixed_err(g["x"], r["x"]) <= TOL
    assert mixed_err(g["res"], r["res"]) <= 1e-6 and (g["res"] < 1e-6).all()
    # iteration budget: MAXITER (2) for the instances that need more, same iterate as the oracle's truncated loop
    g2 = _gpu_sqp(s, batch, max_iter=3)
    r2 = orc.sqp_batch(mirror_opts(opts), batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], max_iter=3)
    assert np.array_equal(g2["status"], r2["status"]) and (r2["status"] == 2).any()
    assert np.array_equal(g2["sqp_iter"], r2["sqp_iter"])
    assert mixed_err(g2["u"], r2["u"]) <= TOL and mixed_err(g2["x"], r2["x"]) <= TOL
    # the plain RTI step still works on the same handle afterwards (finished-instance flags are reset per solve)
    g3 = _gpu_step(s, batch)
    _compare(g3, oracle_batch(mirror_opts(opts), batch))
    s.close()


def test_sqp_mode_with_gp_and_other_qp_kernels(monkeypatch):
    """SQP loop on the GP-augmented model, and through the thread-per-instance / octet QP kernels (skip flags)."""
    B, N = 48, 20
    batch = wl.make_batch(B, N, seed=93, p=1.0, perturb=2.0)
    model = wl.make_gp(M=40, seed=5)
    opts = default_opts(N)
    o = mirror_opts(opts)
    gp = orc.Gp(model)
    gp.apply(o, feat=model["feat"], rows=model["rows"])
    r = orc.sqp_batch(o, batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], gp=gp, gp_state=batch["x0"])
    for v in (7, 4, 1):
        monkeypatch.setenv("ADMPC_QP_VARIANT", str(v))
        s = BatchSolver(B, opts)
        s.set_gp(model)
        s.set_gp_state(batch["x0"])
        g = _gpu_sqp(s, batch)
        assert np.array_equal(g["status"], r["status"]) and np.array_equal(g["sqp_iter"], r["sqp_iter"]), v
        assert mixed_err(g["u"], r["u"]) <= TOL and mixed_err(g["x"], r["x"]) <= TOL, v
        s.close()


def test_acados_shim_sqp_mode():
    """AcadosOcpSolverB200(nlp_solver_type="SQP"): solve() iterates to convergence, get_stats("sqp_iter") reports it."""
    N = 20
    b = wl.make_batch(1, N, seed=95, p=0.0, perturb=2.0)
    opts = default_opts(N)
    cap = AcadosOcpSolverB200(opts, nlp_solver_type="SQP")
    for j in range(N):
        cap.set(j, "yref", b["yref"][0][j * 9:(j + 1) * 9])
        cap.set(j, "p", b["p"][0][j])
    cap.set(N, "yref", b["yref"][0][N * 9:])
    for j in range(N + 1):
        cap.set(j, "x", b["x_init"][0, j])
    cap.set(0, "lbx", b["x0"][0]); cap.set(0, "ubx", b["x0"][0])
    assert cap.solve() == 0
    r = orc.sqp_batch(mirror_opts(opts), b["x0"], b["yref"], b["p"], b["x_init"], b["u_init"])
    assert cap.get_stats("sqp_iter") == r["sqp_iter"][0] > 1
    u = np.stack([cap.get(j, "u") for j in range(N)])
    assert mixed_err(u, r["u"][0]) <= TOL
    cap.free() if hasattr(cap, "free") else None


@pytest.mark.parametrize("B,N,M", [
    (24, 40, 2000),      # BASELINE cfg4 (launch/gp_ad_mpc.launch:6-7 horizon, model_fitting/gp.py:403-430 at M = 2000):
                         # 193 KB GP blob -> one wide CTA per SM in the preparation kernel, two-warp QP kernel
    (48, 20, 450),       # 44 KB blob: the 36..54 KB launch regime of the preparation kernel
    (40, 20, 700),       # 68 KB blob: smallest one-CTA-per-SM case
])
def test_large_gp_model_parity(B, N, M):
    """Every launch regime of the GP preparation kernel on ONE device: linearisation to 1e-10, the whole RTI step to
    1e-8, statuses / QP iteration counts identical."""
    batch = wl.make_batch(B, N, seed=4000 + M, p=1.0)
    rng = np.random.default_rng(M)
    batch["u_init"] = rng.normal(size=(B, N, 2)) * np.array([0.5, 0.1])
    model = wl.make_gp(M=M, seed=9)
    opts = default_opts(N)
    o = mirror_opts(opts)
    gp = orc.Gp(model)
    gp.apply(o, feat=model["feat"], rows=model["rows"])
    s = BatchSolver(B, opts)
    s.set_gp(model)
    g = _gpu_step(s, batch)
    lin = s.get_lin()
    for i in range(0, B, 5):
        it = orc.make_iterate(o, batch["x_init"][i], batch["u_init"][i])
        ref = orc.prepare(o, it, batch["yref"][i], batch["p"][i], gp=gp, gp_state=batch["x0"][i])
        for key in ("A", "B", "b", "q", "r"):
            assert mixed_err(lin[key][i], ref[key]) <= 1e-10, key
    r = oracle_batch(o, batch, gp=gp)
    assert (r["status"] == 0).all()
    _compare(g, r)
    s.close()


@pytest.mark.parametrize("M,N", [(200, 20), (2000, 40)])
def test_gp_fp32_exponent_option_within_stated_bound(M, N):
    """admpc_opts.gp_precision = 1 (opt-in): the RBF exponential runs on the FP32 special-function unit, everything else
    stays FP64.  Stated bound: every kernel value is off by <= 2^-22 relative, so the GP mean is off by
    <= 2^-22 * S, S = sum_i |sigma_f alpha_i k(z, X_i)| (no cancellation credit), and so are, to first order, the entries
    of A, B, b (the GP term enters them times h-weighted O(1) factors).  Checked here: linearisation error <= 2 * 2^-22 * S
    with S evaluated on the batch, statuses and QP iteration counts unchanged, solution within 10x that of the FP64
    oracle.  The default (0) keeps the 1e-10 / 1e-8 bars."""
    B = 32
    batch = wl.make_batch(B, N, seed=5100 + M, p=1.0)
    rng = np.random.default_rng(M + 1)
    batch["u_init"] = rng.normal(size=(B, N, 2)) * np.array([0.5, 0.1])
    model = wl.make_gp(M=M, seed=10)
    opts = default_opts(N, gp_precision=1)
    o = mirror_opts(default_opts(N))
    gp = orc.Gp(model)
    gp.apply(o, feat=model["feat"], rows=model["rows"])
    s = BatchSolver(B, opts)
    s.set_gp(model)
    g = _gpu_step(s, batch)
    lin = s.get_lin()
    worst = 0.0
    for i in range(0, B, 4):
        it = orc.make_iterate(o, batch["x_init"][i], batch["u_init"][i])
        ref = orc.prepare(o, it, batch["yref"][i], batch["p"][i], gp=gp, gp_state=batch["x0"][i])
        for key in ("A", "B", "b"):
            worst = max(worst, mixed_err(lin[key][i], ref[key]))
    # S on the batch: features = states 3..6 of the linearisation points
    z = batch["x_init"][::4, :, 3:7].reshape(-1, 4)
    S = 0.0
    for j in range(model["X"].shape[0]):
        d = (z[:, None, :] - model["X"][j][None, :, :]) / model["ell"][j]
        kk = model["sigma_f"][j] * np.exp(-0.5 * (d * d).sum(axis=2))
        S = max(S, float((np.abs(model["alpha"][j])[None, :] * kk).sum(axis=1).max()))
    bound = 2.0 * 2.0 ** -22 * S
    assert 1e-13 < worst <= bound, (worst, bound)  # really the reduced-precision path, and inside its bound
    assert worst <= 2e-5                           # what the synthetic benchmark models actually show (errors do not add up coherently)
    r = oracle_batch(o, batch, gp=gp)
    assert np.array_equal(g["status"], r["status"]) and np.array_equal(g["qp_iter"], r["qp_iter"])
    eu = max(mixed_err(g["u"], r["u"]), mixed_err(g["x"], r["x"]))
    assert eu <= 10 * bound, (eu, bound)
    print("fp32-exponent GP, M=%d N=%d: S = %.3g, bound %.2e, max lin err %.2e, max solution err %.2e" % (M, N, S, bound, worst, eu))
    s.close()


def test_two_rank_gather_when_two_gpus():
    """Two ranks, two GPUs: the gathered blocks on the root equal the per-rank results for the NCCL send/recv path and for
    the fused peer-memory path (scripts/gather_check.py under torchrun).  Skipped on a one-GPU box."""
    import subprocess, sys
    from ad_mpc_b200 import _lib
    if _lib.load().admpc_device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(root, "scripts", "gather_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    assert "GATHER_OK world=2" in out.stdout and out.stdout.count("FUSED_GATHER_OK") >= 4, out.stdout[-1500:]


def test_two_devices_in_one_process():
    """Handles on different GPUs of one process (per-device kernel attributes, per-handle streams); skipped on 1-GPU boxes."""
    from ad_mpc_b200 import _lib
    if _lib.load().admpc_device_count() < 2:
        pytest.skip("needs two GPUs")
    B, N = 40, 40
    batch = wl.make_batch(B, N, seed=55, p=1.0)
    model = wl.make_gp(M=1500, seed=9)                      # > 48 KB of dynamic shared memory on both devices
    opts = default_opts(N)
    o = mirror_opts(opts)
    gp = orc.Gp(model)
    gp.apply(o, feat=model["feat"], rows=model["rows"])
    r = oracle_batch(o, batch, gp=gp)
    for dev in (0, 1, 0):
        s = BatchSolver(B, opts, device=dev)
        s.set_gp(model)
        g = _gpu_step(s, batch)
        _compare(g, r)
        s.close()
