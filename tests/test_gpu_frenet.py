"""GPU parity of the Frenet-frame model variant (SURVEY 8a A2') against the oracle's Frenet restatement, and of the
dense kernels against the structured Cartesian kernels at zero curvature (where both models coincide)."""
import numpy as np
import pytest

from ad_mpc_b200 import BatchSolver, default_opts, workload as wl
from oracle import oracle as orc
from util_parity import mirror_opts, mixed_err

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _step(s, batch, kappa=None, gp_state=None):
    s.set_iterate(batch["x_init"], batch["u_init"])
    s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"])
    if kappa is not None:
        s.set_kappa(kappa)
    if gp_state is not None:
        s.set_gp_state(gp_state)
    s.solve()
    st, qs, qi = s.get_status()
    return dict(u=s.get_u(), x=s.get_x(), pi=s.get_pi(), status=st, qp_status=qs, qp_iter=qi)


@pytest.mark.parametrize("B,N,p", [(1, 20, 1.0), (70, 20, 1.0), (33, 40, 0.3)])
def test_frenet_rti_step_parity(B, N, p):
    batch = wl.make_batch_frenet(B, N, seed=300 + B, p=p, perturb=2.0)
    opts = default_opts(N, model_variant=1)
    s = BatchSolver(B, opts)
    g = _step(s, batch, kappa=batch["kappa"])
    r = orc.rti_batch(mirror_opts(opts), batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], kappa=batch["kappa"])
    assert np.array_equal(g["status"], r["status"]) and np.array_equal(g["qp_status"], r["qp_status"])
    assert np.array_equal(g["qp_iter"], r["qp_iter"])
    assert mixed_err(g["u"], r["u"]) <= TOL and mixed_err(g["x"], r["x"]) <= TOL
    assert mixed_err(g["pi"], r["pi"]) <= 1e-6
    # second (warm) step from the updated iterate
    s.solve()
    r2 = orc.rti_batch(mirror_opts(opts), batch["x0"], batch["yref"], batch["p"], r["x"], r["u"], kappa=batch["kappa"])
    assert mixed_err(s.get_u(), r2["u"]) <= TOL
    s.close()


def test_frenet_at_zero_curvature_equals_cartesian_kernels():
    """kappa = 0: the dense Frenet kernels must reproduce the structured Cartesian kernels (independent code paths)."""
    B, N = 96, 20
    batch = wl.make_batch(B, N, seed=310, p=0.6, perturb=3.0)
    sc = BatchSolver(B, default_opts(N))
    gc = _step(sc, batch)
    sf = BatchSolver(B, default_opts(N, model_variant=1))
    gf = _step(sf, batch, kappa=np.zeros((B, N)))
    assert np.array_equal(gc["status"], gf["status"]) and np.array_equal(gc["qp_iter"], gf["qp_iter"])
    assert mixed_err(gf["u"], gc["u"]) <= TOL and mixed_err(gf["x"], gc["x"]) <= TOL
    sc.close(); sf.close()


def test_frenet_with_gp_and_nondefault_weights():
    """GP residual on v_y / yaw rate in the Frenet variant + the variant's own weights (q = [0,10,10,10,10,1,0.1],
    r = [10,10], W_e = 0.01 Q: fren_ad_3d_optimizer defaults, SURVEY 8a A2')."""
    B, N = 40, 20
    batch = wl.make_batch_frenet(B, N, seed=320, p=1.0, perturb=1.5)
    q = [0.0, 10.0, 10.0, 10.0, 10.0, 1.0, 0.1]
    opts = default_opts(N, model_variant=1, W=q + [10.0, 10.0], We=[0.01 * v for v in q])
    model = wl.make_gp(M=50, seed=7)
    s = BatchSolver(B, opts)
    s.set_gp(model)
    g = _step(s, batch, kappa=batch["kappa"], gp_state=batch["x0"])
    o = mirror_opts(opts)
    gp = orc.Gp(model)
    gp.apply(o, feat=model["feat"], rows=model["rows"])
    r = orc.rti_batch(o, batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], gp=gp, gp_state=batch["x0"],
                      kappa=batch["kappa"])
    assert np.array_equal(g["status"], r["status"]) and np.array_equal(g["qp_iter"], r["qp_iter"])
    assert mixed_err(g["u"], r["u"]) <= TOL and mixed_err(g["x"], r["x"]) <= TOL
    s.close()


def test_frenet_full_sqp_parity():
    B, N = 40, 20
    batch = wl.make_batch_frenet(B, N, seed=340, p=1.0, perturb=2.5)
    opts = default_opts(N, model_variant=1)
    s = BatchSolver(B, opts)
    s.set_iterate(batch["x_init"], batch["u_init"])
    s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"]); s.set_kappa(batch["kappa"])
    g = s.solve_sqp()
    r = orc.sqp_batch(mirror_opts(opts), batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], kappa=batch["kappa"])
    assert np.array_equal(g["status"], r["status"]) and (r["status"] == 0).all()
    assert np.array_equal(g["sqp_iter"], r["sqp_iter"])
    assert mixed_err(s.get_u(), r["u"]) <= TOL and mixed_err(s.get_x(), r["x"]) <= TOL
    s.close()


def test_frenet_unsupported_calls_fail_loudly():
    from ad_mpc_b200 import _lib
    s = BatchSolver(4, default_opts(20, model_variant=1))
    with pytest.raises(_lib.AdmpcError):
        s.set_track(np.zeros((10, 6)), H=20, traj_dt=0.05)
    sc = BatchSolver(4, default_opts(20))
    with pytest.raises(_lib.AdmpcError):
        sc.set_kappa(np.zeros((4, 20)))
    s.close(); sc.close()


def test_frenet_through_the_acados_shim():
    from ad_mpc_b200 import AcadosOcpSolverB200
    N = 20
    b = wl.make_batch_frenet(1, N, seed=330, p=1.0, perturb=2.0)
    opts = default_opts(N, model_variant=1)
    cap = AcadosOcpSolverB200(opts)
    for j in range(N):
        cap.set(j, "yref", b["yref"][0][j * 9:(j + 1) * 9])
        cap.set(j, "p", b["p"][0][j])
        cap.set(j, "kappa", b["kappa"][0][j])
    cap.set(N, "yref", b["yref"][0][N * 9:])
    for j in range(N + 1):
        cap.set(j, "x", b["x_init"][0, j])
    cap.set(0, "lbx", b["x0"][0]); cap.set(0, "ubx", b["x0"][0])
    assert cap.solve() == 0
    r = orc.rti_batch(mirror_opts(opts), b["x0"], b["yref"], b["p"], b["x_init"], b["u_init"], kappa=b["kappa"])
    u = np.stack([cap.get(j, "u") for j in range(N)])
    assert mixed_err(u, r["u"][0]) <= TOL


def test_frenet_own_constraint_set_through_the_acados_shim(tmp_path):
    """Single-instance shim with con_set = 1: multipliers / slacks in the layout of the reference's iterate dump of this variant
    (ad_mpc/debug.json: 12 rows and 2 + 2 slacks per stage, 20 / 1 + 1 at stage 0), store_iterate / load_iterate round trip."""
    from ad_mpc_b200 import AcadosOcpSolverB200
    N = 20
    b = wl.make_batch_frenet(1, N, seed=331, p=1.0, perturb=1.0)
    b["x_init"][:, :, 1] = np.clip(b["x_init"][:, :, 1], -0.5, 0.5); b["x0"][:, 1] = np.clip(b["x0"][:, 1], -0.5, 0.5)
    opts = _frenet_own_opts(N)

    def fill(cap):
        for j in range(N):
            cap.set(j, "yref", b["yref"][0][j * 9:(j + 1) * 9])
            cap.set(j, "p", b["p"][0][j])
            cap.set(j, "kappa", b["kappa"][0][j])
        cap.set(N, "yref", b["yref"][0][N * 9:])
        cap.set(0, "lbx", b["x0"][0]); cap.set(0, "ubx", b["x0"][0])

    cap = AcadosOcpSolverB200(opts)
    fill(cap)
    for j in range(N + 1):
        cap.set(j, "x", b["x_init"][0, j])
    assert cap.solve() == 0
    r = orc.rti_batch(mirror_opts(opts), b["x0"], b["yref"], b["p"], b["x_init"], b["u_init"], kappa=b["kappa"])
    u = np.stack([cap.get(j, "u") for j in range(N)])
    assert mixed_err(u, r["u"][0]) <= TOL
    assert cap.get(0, "lam").shape == (20,) and cap.get(3, "lam").shape == (12,) and cap.get(0, "sl").shape == (1,) and cap.get(3, "su").shape == (2,)
    Tz = opts.dt * 100.0
    lam3 = cap.get(3, "lam")
    assert abs(lam3[0] + lam3[8] - Tz) < 1e-6 and abs(lam3[3] + lam3[9] - Tz) < 1e-6          # slack stationarity: u0 and delta are the soft rows
    lam0 = cap.get(0, "lam")
    assert abs(lam0[0] + lam0[18] - Tz) < 1e-6 and (lam0[2:9] == 1e-16).all()                # stage 0: u0's slack only, x0 rows eliminated
    # round trip through the acados iterate file: a second solver loaded from it continues exactly like the first one
    f = str(tmp_path / "iterate.json")
    d = cap.store_iterate(f)
    assert len(d["lam_0"]) == 20 and len(d["lam_1"]) == 12 and len(d["sl_0"]) == 1 and len(d["sl_1"]) == 2
    cap2 = AcadosOcpSolverB200(opts)
    fill(cap2)
    cap2.load_iterate(f)
    assert cap.solve() == 0 and cap2.solve() == 0
    for j in (0, 5, N - 1):
        assert np.array_equal(cap.get(j, "u"), cap2.get(j, "u"))
        assert np.allclose(cap.get(j, "lam"), cap2.get(j, "lam"), rtol=0, atol=0)


def test_frenet_kernel_variants_agree(monkeypatch):
    """The warp-per-instance Frenet kernel (6x8 stage structure, default for N <= 63) and the dense thread-per-instance
    kernel (ADMPC_QP_VARIANT=1, also the N > 127 fallback) are independent implementations: same statuses / iteration
    counts, 1e-8 agreement, on a batch with active bounds; and the long-horizon fallback against the oracle."""
    B, N = 64, 20
    batch = wl.make_batch_frenet(B, N, seed=350, p=0.7, perturb=5.0)
    rng = np.random.default_rng(1)
    batch["u_init"] = rng.normal(size=(B, N, 2)) * np.array([1.5, 0.5])
    out = {}
    for v in (4, 1):
        monkeypatch.setenv("ADMPC_QP_VARIANT", str(v))
        s = BatchSolver(B, default_opts(N, model_variant=1, lbu=[-3.0, -1.0], ubu=[2.0, 1.0]))
        out[v] = _step(s, batch, kappa=batch["kappa"])
        s.close()
    assert np.array_equal(out[1]["qp_iter"], out[4]["qp_iter"]) and np.array_equal(out[1]["status"], out[4]["status"])
    assert mixed_err(out[1]["u"], out[4]["u"]) <= TOL and mixed_err(out[1]["x"], out[4]["x"]) <= TOL
    monkeypatch.delenv("ADMPC_QP_VARIANT")
    for Nl in (31, 32, 63, 64, 100, 127, 128):
        bl = wl.make_batch_frenet(10, Nl, seed=360 + Nl, p=1.0, perturb=2.0)
        opts = default_opts(Nl, model_variant=1)
        s = BatchSolver(10, opts)
        g = _step(s, bl, kappa=bl["kappa"])
        r = orc.rti_batch(mirror_opts(opts), bl["x0"], bl["yref"], bl["p"], bl["x_init"], bl["u_init"], kappa=bl["kappa"])
        assert np.array_equal(g["status"], r["status"]) and np.array_equal(g["qp_iter"], r["qp_iter"]), Nl
        assert mixed_err(g["u"], r["u"]) <= TOL and mixed_err(g["x"], r["x"]) <= TOL, Nl
        s.close()


@pytest.mark.parametrize("seed", list(range(6)))
def test_frenet_random_settings(seed):
    """Randomised weights / bounds / horizon / curvature on the Frenet variant (see test_gpu_fuzz.py for the Cartesian twin)."""
    rng = np.random.default_rng(9500 + seed)
    N = int(rng.choice([6, 20, 31, 40, 66]))
    q = list(rng.choice([0.0, 1.0, 10.0, 100.0], size=7))
    kw = dict(dt=float(rng.choice([0.02, 0.05, 0.1])), W=q + list(rng.choice([0.1, 10.0], size=2)), We=[0.01 * v for v in q],
              lbu=[-float(rng.uniform(0.5, 10)), -float(rng.uniform(0.2, 3))], ubu=[float(rng.uniform(0.5, 5)), float(rng.uniform(0.2, 3))],
              lbx=-float(rng.uniform(0.05, 0.6)), ubx=float(rng.uniform(0.05, 0.6)), iter_max=int(rng.choice([3, 50])),
              model_variant=1)
    B = 20
    batch = wl.make_batch_frenet(B, N, seed=200 + seed, dt=kw["dt"], p=float(rng.choice([0.0, 1.0])), perturb=float(rng.choice([1.0, 6.0])),
                                 radius=float(rng.choice([15.0, 50.0, 400.0])))
    batch["u_init"] = rng.normal(size=(B, N, 2)) * np.array([0.5, 0.1])
    opts = default_opts(N, **kw)
    s = BatchSolver(B, opts)
    g = _step(s, batch, kappa=batch["kappa"])
    r = orc.rti_batch(mirror_opts(opts), batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], kappa=batch["kappa"])
    assert np.array_equal(g["status"], r["status"]) and np.array_equal(g["qp_status"], r["qp_status"])
    assert np.array_equal(g["qp_iter"], r["qp_iter"])
    ok = r["status"] == 0
    assert mixed_err(g["u"][ok], r["u"][ok]) <= TOL and mixed_err(g["x"][ok], r["x"][ok]) <= TOL
    s.close()


def test_frenet_through_the_pipelined_host_call():
    """PipelinedSolver (the public batched call with host buffers) on the Frenet variant == resident BatchSolver."""
    from ad_mpc_b200 import PipelinedSolver
    B, N = 300, 20
    batch = wl.make_batch_frenet(B, N, seed=370, p=1.0, perturb=2.0)
    opts = default_opts(N, model_variant=1)
    ps = PipelinedSolver(B, opts, chunks=3)
    ps.set_iterate(batch["x_init"], batch["u_init"]); ps.set_kappa(batch["kappa"])
    u, x, st = np.empty((B, N, 2)), np.empty((B, N + 1, 7)), np.empty(B, dtype=np.int32)
    ps.solve_batch(np.ascontiguousarray(batch["x0"]), np.ascontiguousarray(batch["yref"]), np.ascontiguousarray(batch["p"][:, 0]), u, x, st)
    s = BatchSolver(B, opts)
    g = _step(s, batch, kappa=batch["kappa"])
    assert np.array_equal(st, g["status"]) and np.array_equal(u, g["u"]) and np.array_equal(x, g["x"])
    ps.close(); s.close()


def test_reference_defined_frenet_through_the_pipelined_host_call():
    """The public batched call with host buffers on the variant as the reference defines it (spline + con_set = 1)."""
    from ad_mpc_b200 import PipelinedSolver
    B, N = 200, 20
    batch, breaks, coef = _spline_batch(B, N, 377)
    opts = _frenet_own_opts(N)
    ps = PipelinedSolver(B, opts, chunks=3)
    ps.set_iterate(batch["x_init"], batch["u_init"]); ps.set_kappa_spline(breaks, coef)
    u, x, st = np.empty((B, N, 2)), np.empty((B, N + 1, 7)), np.empty(B, dtype=np.int32)
    ps.solve_batch(np.ascontiguousarray(batch["x0"]), np.ascontiguousarray(batch["yref"]), np.ascontiguousarray(batch["p"][:, 0]), u, x, st)
    s = BatchSolver(B, opts)
    s.set_kappa_spline(breaks, coef)
    g = _step(s, batch)
    assert (g["status"] == 0).any()
    assert np.array_equal(st, g["status"]) and np.array_equal(u, g["u"]) and np.array_equal(x, g["x"])
    ps.close(); s.close()


def _spline_batch(B, N, seed):
    """Curvature that really varies along the horizon: kappa(s) = 0.02 + 0.012 sin(s / 9 + phase_i), sampled at knots
    every 4 m around each vehicle and turned into the not-a-knot cubic the reference's bspline interpolant builds."""
    from ad_mpc_b200 import kappa_pp_from_knots
    batch = wl.make_batch_frenet(B, N, seed=seed, p=1.0, perturb=2.0)
    rng = np.random.default_rng(seed + 1)
    phase = rng.uniform(0, 6.28, size=B)
    K = 12
    breaks, coef = np.zeros((B, K + 1)), np.zeros((B, K, 4))
    for i in range(B):
        kn = batch["x0"][i, 0] - 6.0 + 4.0 * np.arange(K + 1)
        breaks[i], coef[i] = kappa_pp_from_knots(kn, 0.02 + 0.012 * np.sin(kn / 9.0 + phase[i]))
    return batch, breaks, coef


@pytest.mark.parametrize("B,N", [(48, 20), (20, 40)])
def test_frenet_spline_curvature_inside_the_model(B, N):
    """kappa(s) as a spline evaluated inside the model at every RK4 sub-stage, with the d kappa / d s column of the
    Jacobian (the reference's bspline-interpolant semantics): linearisation and RTI step against the oracle."""
    batch, breaks, coef = _spline_batch(B, N, 420 + N)
    opts = default_opts(N, model_variant=1)
    o = mirror_opts(opts)
    s = BatchSolver(B, opts)
    s.set_kappa_spline(breaks, coef)
    g = _step(s, batch)
    orc.set_batch_kappa_spline(breaks, coef)
    try:
        r = orc.rti_batch(o, batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], kappa=np.zeros((B, N)))
    finally:
        orc.set_batch_kappa_spline(None)
    assert np.array_equal(g["status"], r["status"]) and np.array_equal(g["qp_iter"], r["qp_iter"])
    assert mixed_err(g["u"], r["u"]) <= TOL and mixed_err(g["x"], r["x"]) <= TOL
    # the column of s is really there: the same spline frozen at the nodes (d kappa / d s dropped) gives another answer
    kap_nodes = np.zeros((B, N))
    for i in range(B):
        for k in range(N):
            sk = batch["x_init"][i, k, 0]
            j = int(np.clip(np.searchsorted(breaks[i], sk, side="right") - 1, 0, coef.shape[1] - 1))
            t = sk - breaks[i, j]
            kap_nodes[i, k] = coef[i, j] @ np.array([1.0, t, t * t, t ** 3])
    s.set_kappa_spline(None)
    g2 = _step(s, batch, kappa=kap_nodes)
    assert mixed_err(g2["u"], g["u"]) > 1e-7
    # constant spline == per-node constant (both Jacobian forms coincide)
    cb = np.tile(np.array([-1e3, 1e3]), (B, 1)); cc = np.zeros((B, 1, 4)); cc[:, 0, 0] = 0.02
    s.set_kappa_spline(cb, cc)
    g3 = _step(s, batch)
    s.set_kappa_spline(None)
    g4 = _step(s, batch, kappa=np.full((B, N), 0.02))
    assert np.array_equal(g3["qp_iter"], g4["qp_iter"]) and mixed_err(g3["u"], g4["u"]) <= TOL
    s.close()


def _frenet_own_opts(N, **kw):
    """The Frenet variant's own OCP (SURVEY 8a A2' + ad_mpc/debug.json): weights q = [0,10,10,10,10,1,0.1], r = [10,10],
    W_e = 0.01 Q, steering rate in [-2, 2] hard, acceleration soft, e_y in [-2, 2] hard, steering angle soft, zl = zu = 100."""
    q = [0.0, 10.0, 10.0, 10.0, 10.0, 1.0, 0.1]
    return default_opts(N, model_variant=1, con_set=1, W=q + [10.0, 10.0], We=[0.01 * v for v in q], zl=[100.0, 100.0],
                        zu=[100.0, 100.0], lbu=[-10.0, -2.0], ubu=[5.0, 2.0], lbx=-0.52, ubx=0.52, lbx2=-2.0, ubx2=2.0, **kw)


@pytest.mark.parametrize("B,N,tight", [(48, 20, False), (40, 20, True), (17, 40, True)])
def test_frenet_own_constraint_set_parity(B, N, tight):
    """con_set = 1: u0 soft, u1 hard, e_y hard, steering angle soft (12 rows per stage), against the oracle's generic
    constraint-set IPM; `tight` shrinks the boxes so that the hard e_y bound, the hard steering-rate bound and the soft
    steering-angle bound are all active somewhere in the batch."""
    batch = wl.make_batch_frenet(B, N, seed=360 + B, p=1.0, perturb=3.0)
    kw = dict(lbx2=-0.6, ubx2=0.6, lbu=[-2.0, -0.5], ubu=[1.5, 0.5], lbx=-0.08, ubx=0.08) if tight else {}
    opts = _frenet_own_opts(N)
    for k, v in kw.items():
        cur = getattr(opts, k)
        if hasattr(cur, "__len__"):
            for i, vi in enumerate(v):
                cur[i] = vi
        else:
            setattr(opts, k, v)
    if tight:                                           # iterate inside the hard boxes (an interior point exists)
        batch["x_init"][:, :, 1] = np.clip(batch["x_init"][:, :, 1], -0.5, 0.5)
        batch["x0"][:, 1] = np.clip(batch["x0"][:, 1], -0.5, 0.5)
    s = BatchSolver(B, opts)
    g = _step(s, batch, kappa=batch["kappa"])
    r = orc.rti_batch(mirror_opts(opts), batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], kappa=batch["kappa"])
    # An instance whose hard e_y box is infeasible (x0 outside it) never converges: its IPM ends in a minimum-step failure
    # (status 4) or runs into the iteration limit (tolerated by RTI, status 0), and which of the two is rounding-dependent (the
    # -DADMPC_DEBUG build takes the other branch than the release build on one instance of this batch).  Everything is compared
    # strictly where the oracle's QP converged; elsewhere the GPU must not claim convergence either.
    solid = r["qp_status"] == 0
    assert solid.sum() >= B // 2
    assert np.array_equal(g["status"][solid], r["status"][solid]) and np.array_equal(g["qp_status"][solid], r["qp_status"][solid])
    assert (g["qp_status"][~solid] != 0).all()
    good = solid & (r["status"] == 0) & (g["status"] == 0)
    assert np.array_equal(g["qp_iter"][good], r["qp_iter"][good])
    same = g["status"] == r["status"]                       # (a failed QP leaves the iterate untouched, a tolerated one updates it)
    assert mixed_err(g["u"][same], r["u"][same]) <= TOL and mixed_err(g["x"][same], r["x"][same]) <= TOL
    assert mixed_err(g["pi"][good], r["pi"][good]) <= 1e-6
    lam, t = s.get_lam(), s.get_t()
    assert lam.shape == (B, N, 12) and t.shape == (B, N, 12)
    ok = (g["status"] == 0) & (g["qp_status"] == 0)
    assert ok.any()
    if tight:
        # hard rows active: steering rate (rows 1 / 5) and e_y (rows 2 / 6) ; soft steering angle: lam_ls1 + lam_lbx_delta = Ts zl
        assert (lam[ok][:, :, [1, 5]] > 1e-3).any() and (lam[ok][:, 1:, [2, 6]] > 1e-3).any()
        assert (lam[ok][:, 1:, [3, 7]] > 1e-3).any()
    Tz = opts.dt * 100.0
    assert np.abs(lam[ok][:, 1:, 3] + lam[ok][:, 1:, 9] - Tz).max() < 1e-6      # slack stationarity of the soft steering angle
    assert np.abs(lam[ok][:, :, 0] + lam[ok][:, :, 8] - Tz).max() < 1e-6       # ... of the soft acceleration
    # second (warm) step from the updated iterate
    s.solve()
    r2 = orc.rti_batch(mirror_opts(opts), batch["x0"], batch["yref"], batch["p"], r["x"], r["u"], kappa=batch["kappa"])
    solid2 = same & (r2["qp_status"] == 0)
    assert np.array_equal(s.get_status()[0][solid2], r2["status"][solid2])
    assert mixed_err(s.get_u()[solid2], r2["u"][solid2]) <= TOL
    s.close()


def test_frenet_own_constraint_set_full_sqp_and_spline_curvature():
    """The variant as the reference defines it end to end: kappa(s) spline inside the model + its own constraint set, SQP mode."""
    from ad_mpc_b200.solver import kappa_pp_from_knots
    B, N = 24, 20
    batch = wl.make_batch_frenet(B, N, seed=371, p=1.0, perturb=2.0)
    opts = _frenet_own_opts(N)
    s_knots = np.linspace(-50.0, 450.0, 26)
    rng = np.random.default_rng(5)
    brk, cf = [], []
    for b in range(B):
        bb, cc = kappa_pp_from_knots(s_knots, 0.02 + 0.01 * np.sin(0.03 * s_knots + rng.uniform(0, 6.28)))
        brk.append(bb); cf.append(cc)
    brk, cf = np.stack(brk), np.stack(cf)
    s = BatchSolver(B, opts)
    s.set_iterate(batch["x_init"], batch["u_init"])
    s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"]); s.set_kappa(batch["kappa"])
    s.set_kappa_spline(brk, cf)
    g = s.solve_sqp()
    orc.set_batch_kappa_spline(brk, cf)
    try:
        r = orc.sqp_batch(mirror_opts(opts), batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], kappa=batch["kappa"])
    finally:
        orc.set_batch_kappa_spline(None)
    assert np.array_equal(g["status"], r["status"]) and np.array_equal(g["sqp_iter"], r["sqp_iter"])
    assert mixed_err(s.get_u(), r["u"]) <= TOL and mixed_err(s.get_x(), r["x"]) <= TOL
    s.close()


@pytest.mark.parametrize("N,own_set,spline", [(20, True, True), (20, False, True), (20, True, False), (40, True, True)])
def test_frenet_generic_tensor_core_kernel_agrees_with_the_dense_kernel(monkeypatch, N, own_set, spline):
    """qp_mma_g (dense column of s and / or con_set = 1 on the FP64 tensor cores, the default for N <= 63) against the dense
    thread-per-instance kernel (ADMPC_QP_VARIANT=1): independent implementations, same statuses / iteration counts, 1e-8
    on the trajectories and 1e-6 on the multipliers; the launch count proves which kernel ran (fused update: 2 per step)."""
    B = 40
    batch, breaks, coef = _spline_batch(B, N, 500 + N)
    opts = _frenet_own_opts(N) if own_set else default_opts(N, model_variant=1, lbu=[-3.0, -1.0], ubu=[2.0, 1.0])
    if own_set:                                         # boxes tight enough that hard and soft rows are active
        opts.lbx2, opts.ubx2, opts.lbx, opts.ubx = -0.6, 0.6, -0.08, 0.08
        opts.lbu[0], opts.lbu[1], opts.ubu[0], opts.ubu[1] = -2.0, -0.5, 1.5, 0.5
        batch["x_init"][:, :, 1] = np.clip(batch["x_init"][:, :, 1], -0.5, 0.5)
        batch["x0"][:, 1] = np.clip(batch["x0"][:, 1], -0.5, 0.5)
    out, launches = {}, {}
    for v in (0, 1):
        if v:
            monkeypatch.setenv("ADMPC_QP_VARIANT", str(v))
        s = BatchSolver(B, opts)
        if spline:
            s.set_kappa_spline(breaks, coef)
        n0 = s.kernel_launches()
        out[v] = _step(s, batch, kappa=None if spline else batch["kappa"])
        launches[v] = s.kernel_launches() - n0
        out[v]["lam"], out[v]["t"] = s.get_lam(), s.get_t()
        out[v]["sl"], out[v]["su"] = s.get_slacks()
        s.close()
    monkeypatch.delenv("ADMPC_QP_VARIANT")
    assert launches[1] - launches[0] == 1               # prepare + fused QP / update  vs  prepare + dense QP + update
    a, b = out[0], out[1]
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["qp_status"], b["qp_status"])
    good = b["status"] == 0
    assert good.any() and np.array_equal(a["qp_iter"][good], b["qp_iter"][good])
    assert mixed_err(a["u"], b["u"]) <= TOL and mixed_err(a["x"], b["x"]) <= TOL
    for f in ("pi", "lam", "t", "sl", "su"):
        assert mixed_err(a[f][good], b[f][good]) <= 1e-6, f


@pytest.mark.parametrize("seed", list(range(8)))
def test_frenet_reference_defined_variant_random_settings(seed):
    """Randomised weights / bounds / horizon / GP on the variant as the reference defines it -- kappa(s) spline inside the model,
    own constraint set (con_set = 1) -- through the default kernels (gp_sweep_kernel<FR>, prepare_dense_kernel, qp_mma_g_kernel)
    against the oracle: statuses, QP statuses, iteration counts of the solved instances, 1e-8 on the trajectories."""
    rng = np.random.default_rng(9700 + seed)
    N = int(rng.choice([5, 20, 31, 32, 48, 63]))
    B = 24
    q = [0.0] + list(rng.choice([0.1, 1.0, 10.0, 100.0], size=6))
    opts = _frenet_own_opts(N, dt=float(rng.choice([0.02, 0.05, 0.1])), iter_max=int(rng.choice([4, 50])))
    for i, v in enumerate(q + list(rng.choice([0.1, 10.0], size=2))):
        opts.W[i] = v
    for i, v in enumerate(q):
        opts.We[i] = 0.01 * v
    opts.lbu[0], opts.ubu[0] = -float(rng.uniform(0.5, 10)), float(rng.uniform(0.5, 5))
    opts.lbu[1], opts.ubu[1] = -float(rng.uniform(0.3, 3)), float(rng.uniform(0.3, 3))
    opts.lbx, opts.ubx = -float(rng.uniform(0.05, 0.6)), float(rng.uniform(0.05, 0.6))
    opts.lbx2, opts.ubx2 = -float(rng.uniform(0.6, 3)), float(rng.uniform(0.6, 3))
    for j in range(2):
        opts.zl[j] = opts.zu[j] = float(rng.choice([10.0, 100.0, 1000.0]))
    batch, breaks, coef = _spline_batch(B, N, 600 + seed)
    batch["x_init"][:, :, 1] = np.clip(batch["x_init"][:, :, 1], -0.5, 0.5)      # inside the hard e_y box
    batch["x0"][:, 1] = np.clip(batch["x0"][:, 1], -0.5, 0.5)
    batch["u_init"] = rng.normal(size=(B, N, 2)) * np.array([0.5, 0.1])
    s = BatchSolver(B, opts)
    o = mirror_opts(opts)
    gp = None
    if seed % 2:
        model = wl.make_gp(M=int(rng.choice([16, 100])), seed=30 + seed)
        s.set_gp(model)
        gp = orc.Gp(model)
        gp.apply(o, feat=model["feat"], rows=model["rows"])
    s.set_kappa_spline(breaks, coef)
    g = _step(s, batch)
    orc.set_batch_kappa_spline(breaks, coef)
    try:
        r = orc.rti_batch(o, batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], kappa=np.zeros((B, N)), gp=gp)
    finally:
        orc.set_batch_kappa_spline(None)
    assert np.array_equal(g["status"], r["status"]) and np.array_equal(g["qp_status"], r["qp_status"])
    ok = r["status"] == 0
    assert ok.any()
    conv = ok & (r["qp_status"] == 0)
    assert np.array_equal(g["qp_iter"][conv], r["qp_iter"][conv])
    assert mixed_err(g["u"][ok], r["u"][ok]) <= TOL and mixed_err(g["x"][ok], r["x"][ok]) <= TOL
    s.close()


def test_reference_defined_frenet_full_size(monkeypatch):
    """BASELINE cfg-3 size (B = 16384, N = 20, GP M = 200) on the variant as the reference defines it: size-independent properties
    instead of a 16k-instance oracle run -- (i) duplicated instances give bit-identical results wherever they sit in the batch,
    (ii) a 128-instance sample matches the oracle, (iii) the dense thread-per-instance kernel (independent implementation) agrees
    on the whole batch, (iv) every status is 0."""
    B, N, S = 16384, 20, 128
    base, breaks, coef = _spline_batch(S, N, 731)
    base["x_init"][:, :, 1] = np.clip(base["x_init"][:, :, 1], -0.5, 0.5)
    base["x0"][:, 1] = np.clip(base["x0"][:, 1], -0.5, 0.5)
    reps = B // S
    batch = {k: np.concatenate([v] * reps, axis=0) for k, v in base.items()}
    bk, cf = np.concatenate([breaks] * reps, axis=0), np.concatenate([coef] * reps, axis=0)
    model = wl.make_gp(M=200, seed=20263)
    opts = _frenet_own_opts(N)
    out = {}
    for v in (0, 1):
        if v:
            monkeypatch.setenv("ADMPC_QP_VARIANT", "1")
        s = BatchSolver(B, opts)
        s.set_gp(model)
        s.set_kappa_spline(bk, cf)
        out[v] = _step(s, batch)
        s.close()
    monkeypatch.delenv("ADMPC_QP_VARIANT")
    g = out[0]
    assert (g["status"] == 0).all() and (g["qp_status"] == 0).all()
    u = g["u"].reshape(reps, S, N, 2)
    assert np.array_equal(u[0], u[-1]) and np.array_equal(u[0], u[reps // 2])
    assert np.array_equal(g["qp_iter"], out[1]["qp_iter"]) and np.array_equal(g["status"], out[1]["status"])
    assert mixed_err(g["u"], out[1]["u"]) <= TOL and mixed_err(g["x"], out[1]["x"]) <= TOL
    o = mirror_opts(opts)
    gp = orc.Gp(model)
    gp.apply(o, feat=model["feat"], rows=model["rows"])
    orc.set_batch_kappa_spline(breaks, coef)
    try:
        r = orc.rti_batch(o, base["x0"], base["yref"], base["p"], base["x_init"], base["u_init"], kappa=np.zeros((S, N)), gp=gp)
    finally:
        orc.set_batch_kappa_spline(None)
    assert np.array_equal(g["qp_iter"][:S], r["qp_iter"]) and np.array_equal(g["status"][:S], r["status"])
    assert mixed_err(u[0], r["u"]) <= TOL and mixed_err(g["x"][:S], r["x"]) <= TOL


def test_con_set_1_needs_the_frenet_model():
    from ad_mpc_b200 import _lib
    with pytest.raises(_lib.AdmpcError):
        BatchSolver(4, default_opts(20, con_set=1))
