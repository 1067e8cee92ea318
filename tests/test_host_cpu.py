"""CPU-only tests (-m "not gpu"): C-ABI surface, host logic, oracle self-consistency, world_size-2 sharding."""
import ctypes as C
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "admpc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:admpc|sim_car_acados)_\w+)\s*\(", txt)))


def test_cabi_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from ad_mpc_b200 import _lib
    L = C.CDLL(_lib.SO_PATH)
    names = _header_functions()
    assert len(names) >= 50
    for n in names:
        assert hasattr(L, n), "libadmpc_b200.so does not export " + n
        assert n in _lib.SYMBOLS, "ctypes binding misses " + n
    assert sorted(_lib.SYMBOLS) == names


def test_opts_struct_layout_matches_header():
    """ctypes mirror of admpc_opts must have the C struct's size (guards against silent ABI drift)."""
    from ad_mpc_b200 import _lib
    src = '#include <stdio.h>\n#include "admpc.h"\nint main(){printf("%zu %zu\\n", sizeof(admpc_opts), __builtin_offsetof(admpc_opts, dt));return 0;}'
    exe = os.path.join(ROOT, "tests", "_sizeof_opts")
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe], input=src.encode(), check=True)
    size, off = map(int, subprocess.check_output([exe]).split())
    os.remove(exe)
    assert size == C.sizeof(_lib.AdmpcOpts) and off == _lib.AdmpcOpts.dt.offset


def test_product_fails_loudly_without_gpu():
    from ad_mpc_b200 import _lib, BatchSolver
    L = _lib.load()
    if L.admpc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.AdmpcError):
        BatchSolver(8)
    cap = L.sim_car_acados_create_capsule()
    assert L.sim_car_acados_create(cap) == -2            # ADMPC_E_CUDA: no CPU fallback
    assert L.sim_car_acados_solve(cap) == -3             # ADMPC_E_STATE: never created
    L.sim_car_acados_free_capsule(cap)


def test_default_opts_match_reference_constants():
    """acados_solver_sim_car.c:362,393-399,481-485,455-473,549-552,595-596 and ad_3d.py:47-60."""
    from ad_mpc_b200 import default_opts
    o = default_opts(20)
    assert list(o.W) == [10, 10, 100, 0, 0, 0, 0, 1, 100] and list(o.We) == [1e-5, 1e-5, 1e-4, 0, 0, 0, 0]
    assert o.dt == 0.05 and list(o.lbu) == [-10, -3] and list(o.ubu) == [5, 3] and (o.lbx, o.ubx) == (-0.52, 0.52)
    assert list(o.zl) == [10, 10] and list(o.Zu) == [0, 0] and o.iter_max == 50
    assert abs(o.lf - 1.08) < 1e-12 and abs(o.lr - 1.62) < 1e-12 and abs(o.iz - 2624.4) < 1e-9 and o.mass == 1500
    assert abs(o.cf2 - 83458.139053772335) < 1e-6 and abs(o.cr2 - 55638.759369181564) < 1e-6   # literals in sim_car_expl_ode_fun.c


def test_oracle_and_product_defaults_agree():
    from ad_mpc_b200 import default_opts
    from oracle import oracle as orc
    from util_parity import OPT_FIELDS
    a, b = default_opts(20), orc.default_opts(20)
    for f in OPT_FIELDS:
        va, vb = getattr(a, f), getattr(b, f)
        assert (list(va) == list(vb)) if hasattr(va, "__len__") else (va == vb), f


def test_unwrap_matches_reference_loop():
    """workload.unwrap_ref_psi == literal restatement of ad_3d_optimizer.py:423-437."""
    from ad_mpc_b200 import workload as wl
    rng = np.random.default_rng(0)
    psi0 = rng.uniform(-math.pi, math.pi, size=200)
    ref = rng.uniform(-math.pi, math.pi, size=(200, 21))
    out = wl.unwrap_ref_psi(psi0, ref)
    for i in range(200):
        for j in range(21):
            r = ref[i, j]
            if psi0[i] < 0:
                if psi0[i] + math.pi < r:
                    r = r - 2 * math.pi
            elif psi0[i] > 0:
                if psi0[i] - math.pi > r:
                    r = r + 2 * math.pi
            assert out[i, j] == r


def test_oracle_gp_matches_numpy_restatement_and_fd():
    """orc_gp_predict == gp.py:426-430 (cdist / exp / dot) and its gradient == central differences."""
    from scipy.spatial.distance import cdist
    from ad_mpc_b200 import workload as wl
    from oracle import oracle as orc
    model = wl.make_gp(M=80, seed=5)
    o = orc.default_opts(20)
    gp = orc.Gp(model)
    gp.apply(o, feat=model["feat"], rows=model["rows"])
    rng = np.random.default_rng(1)
    for _ in range(20):
        z = rng.uniform(wl.GP_BOX_LO, wl.GP_BOX_HI)
        mu, dmu = orc.gp_predict(o, gp, z)
        for j in range(2):
            ell = model["ell"][j]
            k_s = model["sigma_f"][j] * np.exp(-0.5 * cdist(z[None] / ell, model["X"][j] / ell, metric="sqeuclidean"))
            ref = (k_s @ model["alpha"][j])[0] + model["y_mean"][j]
            assert abs(mu[j] - ref) <= 1e-9 * max(1.0, abs(ref))
        h = 1e-6
        for d in range(4):
            zp, zm = z.copy(), z.copy()
            zp[d] += h
            zm[d] -= h
            fd = (orc.gp_predict(o, gp, zp)[0] - orc.gp_predict(o, gp, zm)[0]) / (2 * h)
            assert np.abs(fd - dmu[:, d]).max() < 1e-5 * max(1.0, np.abs(dmu[:, d]).max())


def test_oracle_jacobian_fd_with_gp():
    from ad_mpc_b200 import workload as wl
    from oracle import oracle as orc
    model = wl.make_gp(M=40, seed=6)
    o = orc.default_opts(20)
    gp = orc.Gp(model)
    gp.apply(o, feat=model["feat"], rows=model["rows"], stage0_trigger=0)
    rng = np.random.default_rng(2)
    x = np.array([1.0, 2.0, 0.3, 8.0, 0.2, 0.1, 0.05])
    u = np.array([0.5, -0.1])
    for p in (0.0, 1.0):
        f, Jx, Ju = orc.model_jac(o, x, u, p, gp=gp)
        h = 1e-6
        for j in range(7):
            xp, xm = x.copy(), x.copy()
            xp[j] += h
            xm[j] -= h
            fd = (orc.model_jac(o, xp, u, p, gp=gp)[0] - orc.model_jac(o, xm, u, p, gp=gp)[0]) / (2 * h)
            assert np.abs(fd - Jx[:, j]).max() < 2e-5 * max(1.0, np.abs(Jx[:, j]).max())
        for j in range(2):
            up, um = u.copy(), u.copy()
            up[j] += h
            um[j] -= h
            fd = (orc.model_jac(o, x, up, p, gp=gp)[0] - orc.model_jac(o, x, um, p, gp=gp)[0]) / (2 * h)
            assert np.abs(fd - Ju[:, j]).max() < 2e-5


def test_oracle_qp_solution_satisfies_kkt_independently():
    """The IPM's answer is checked against the QP's KKT conditions assembled here in numpy (not by the oracle):
    convex QP => KKT-feasible point == the unique minimiser.  Uses a batch with active input/steering bounds."""
    from ad_mpc_b200 import workload as wl
    from oracle import oracle as orc
    N = 20
    o = orc.default_opts(N)
    batch = wl.make_batch(24, N, seed=99, perturb=6.0)
    Ts, W, We = o.dt, np.array(o.W[:]), np.array(o.We[:])
    n_active = 0
    for i in range(24):
        it = orc.make_iterate(o, batch["x_init"][i], batch["u_init"][i])
        lin = orc.prepare(o, it, batch["yref"][i], batch["p"][i])
        sol = orc.qp_solve(o, lin["_c"], it, batch["x0"][i])
        assert sol["qp_status"] == 0
        dx, du, pi, lam, t, sl, su = (sol[k] for k in ("dx", "du", "pi", "lam", "t", "sl", "su"))
        xbar, ubar = batch["x_init"][i], batch["u_init"][i]
        assert np.abs(dx[0] - (batch["x0"][i] - xbar[0])).max() < 1e-12
        for k in range(N):
            A, B = lin["A"][k], lin["B"][k]
            assert np.abs(A @ dx[k] + B @ du[k] + lin["b"][k] - dx[k + 1]).max() < 1e-7
            gu = Ts * W[7:] * du[k] + lin["r"][k] + B.T @ pi[k] - lam[k][0:2] + lam[k][3:5]
            assert np.abs(gu).max() < 1e-7
            assert np.abs(Ts * 10 - lam[k][0:2] - lam[k][6:8]).max() < 1e-7
            assert np.abs(Ts * 10 - lam[k][3:5] - lam[k][8:10]).max() < 1e-7
            lo, hi = np.array(o.lbu[:]) - ubar[k], np.array(o.ubu[:]) - ubar[k]
            assert (du[k] - lo + sl[k] >= -1e-7).all() and (hi - du[k] + su[k] >= -1e-7).all()
            assert (sl[k] >= -1e-12).all() and (su[k] >= -1e-12).all()
            if k >= 1:
                gx = Ts * W[:7] * dx[k] + lin["q"][k] + A.T @ pi[k] - pi[k - 1]
                gx[6] += -lam[k][2] + lam[k][5]
                assert np.abs(gx).max() < 1e-7
                d6 = xbar[k][6] + dx[k][6]
                assert o.lbx - 1e-7 <= d6 <= o.ubx + 1e-7
            on = np.ones(10, bool)
            if k == 0:
                on[[2, 5]] = False
            assert (lam[k][on] >= 0).all() and (t[k][on] > 0).all() and (lam[k][on] * t[k][on]).max() < 1e-7
            n_active += int((lam[k][[0, 1, 3, 4]] > 1e-3).sum())
        gN = We * dx[N] + lin["q"][N] - pi[N - 1]
        assert np.abs(gN).max() < 1e-7
    assert n_active > 10, "batch does not exercise active bounds"


def test_shard_ranges_cover_batch():
    from ad_mpc_b200.shard import gather_order, shard_range
    for B in (1, 7, 16384, 131072, 131073):
        for world in (1, 2, 3, 4, 8):
            rs = gather_order(B, world)
            assert rs[0][0] == 0 and rs[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from ad_mpc_b200 import workload as wl
from ad_mpc_b200.shard import shard_range, gather_order
from oracle import oracle as orc
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
B, N = 37, 20
batch = wl.make_batch(B, N, seed=3)
model = [wl.make_gp(M=16, seed=4) if rank == 0 else None]
dist.broadcast_object_list(model, src=0)                 # "GP-model broadcast"
model = model[0]
o = orc.default_opts(N)
gp = orc.Gp(model); gp.apply(o, feat=model["feat"], rows=model["rows"])
lo, hi = shard_range(B, rank, world)
r = orc.rti_batch(o, batch["x0"][lo:hi], batch["yref"][lo:hi], batch["p"][lo:hi], batch["x_init"][lo:hi], batch["u_init"][lo:hi], gp=gp)
blocks = [None] * world if rank == 0 else None
dist.gather_object((r["u"], r["x"], r["status"]), blocks, dst=0)   # "solution gather"
if rank == 0:
    full = orc.rti_batch(o, batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], gp=gp)
    u = np.concatenate([b[0] for b in blocks]); x = np.concatenate([b[1] for b in blocks]); st = np.concatenate([b[2] for b in blocks])
    assert [ (hi_-lo_) for lo_,hi_ in gather_order(B, world)] == [b[0].shape[0] for b in blocks]
    assert np.array_equal(u, full["u"]) and np.array_equal(x, full["x"]) and np.array_equal(st, full["status"])
    print("SHARD_OK")
dist.destroy_process_group()
'''


def test_world_size_2_shard_broadcast_gather_gloo(tmp_path):
    """N>1 host path on CPU: block sharding + model broadcast + gather to rank 0 reproduce the single-process batch
    (gloo, world_size 2; the oracle stands in for the CUDA solver, which is allowed in tests only)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "SHARD_OK" in out.stdout


def test_c_example_compiles_against_the_header():
    """examples/main_sim_car_b200.c (the reference's main_sim_car.c call sequence) builds with plain gcc + the C ABI."""
    import __graft_entry__ as g
    g.build()
    exe = os.path.join(ROOT, "examples", "main_sim_car_b200")
    subprocess.run(["/usr/bin/gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "main_sim_car_b200.c"),
                    "-o", exe, "-L", os.path.join(ROOT, "ad_mpc_b200"), "-ladmpc_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "ad_mpc_b200"), "-lm"], check=True)
    assert os.path.exists(exe)


def test_optimizer_marshalling_matches_the_reference_loops():
    """AD3DOptimizerB200.marshal (vectorised) against the reference's per-node Python of run_optimization
    (ad_3d_optimizer.py:417-443): heading unwrap relative to x0.psi on every node, in-place edit of the terminal
    target, vel_switch blend."""
    import math
    from ad_mpc_b200.optimizer import AD3DOptimizerB200
    rng = np.random.default_rng(11)
    B, N = 64, 20
    x_init = rng.normal(size=(B, 7))
    x_init[:, 2] = rng.uniform(-math.pi, math.pi, size=B)
    x_init[:8, 2] = 0.0                                            # psi == 0: neither branch
    x_init[:, 3] = rng.uniform(0, 14, size=B)
    target = rng.normal(size=(B, N + 1, 7))
    target[..., 2] = rng.uniform(-math.pi, math.pi, size=(B, N + 1))
    u_target = rng.normal(size=(B, N + 1, 2))
    t_in = target.copy()
    yref, vs = AD3DOptimizerB200.marshal(x_init, target, u_target, N, 3.0, 12.0)
    for b in range(B):
        stacked = t_in[b].copy()
        for j in range(N):
            ref = np.concatenate((stacked[j, :], u_target[b, j, :]))
            if x_init[b, 2] < 0:
                if x_init[b, 2] + math.pi < ref[2]:
                    ref[2] = ref[2] - 2 * math.pi
            elif x_init[b, 2] > 0:
                if x_init[b, 2] - math.pi > ref[2]:
                    ref[2] = ref[2] + 2 * math.pi
            assert np.array_equal(yref[b, j * 9:(j + 1) * 9], ref)
        if x_init[b, 2] < 0:
            if x_init[b, 2] + math.pi < stacked[N, 2]:
                stacked[N, 2] = stacked[N, 2] - 2 * math.pi
        elif x_init[b, 2] > 0:
            if x_init[b, 2] - math.pi > stacked[N, 2]:
                stacked[N, 2] = stacked[N, 2] + 2 * math.pi
        assert np.array_equal(yref[b, N * 9:], stacked[N, :]) and np.array_equal(target[b], stacked)
        assert vs[b] == min(max((x_init[b, 3] - 3.0) / (12.0 - 3.0), 0.0), 1.0)
    # backup control slice arithmetic of :472
    prev = np.arange(40.0)
    w = AD3DOptimizerB200._backup(prev)
    assert w.shape == (40,) and np.array_equal(w[:39], np.concatenate((prev[2:-1], prev[-3:-1])))


def test_oracle_defaults_equal_product_defaults():
    """bench.py's reference arm builds its options from the oracle's own defaults (it must not import the product):
    they have to be the product's defaults, field by field."""
    from ad_mpc_b200 import default_opts
    from oracle import oracle as orc
    from util_parity import OPT_FIELDS
    for N in (20, 40):
        po, oo = default_opts(N), orc.default_opts(N=N)
        for f in OPT_FIELDS:
            a, b = getattr(po, f), getattr(oo, f)
            if hasattr(a, "__len__"):
                assert list(a) == list(b)[:len(a)], f
            else:
                assert a == b, f


def test_reference_arm_never_loads_the_product():
    """`bench.py --impl reference` runs the CPU restatement only: no ad_mpc_b200 import, no libadmpc_b200.so in the process."""
    code = ("import sys, json, io, contextlib; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0'];"
            "import bench; bench.CFG['B']=64;"
            "buf=io.StringIO();\n"
            "with contextlib.redirect_stdout(buf): bench.main()\n"
            "line=json.loads(buf.getvalue().strip().splitlines()[-1]);"
            "assert line['impl']=='reference' and line['value']>0 and 'reference_note' not in line['config'];"
            "assert 'ad_mpc_b200' not in sys.modules;"
            "assert 'libadmpc_b200' not in open('/proc/self/maps').read(); print('OK')")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-500:] + out.stderr[-1500:]
