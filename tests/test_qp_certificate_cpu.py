"""Solver-independent certificate for the QP solutions of the oracle's interior-point method.

The IPM iterate path cannot be pinned (HPIPM is not in the reference tree), but the QP is strictly convex in the inputs
and slacks, so its SOLUTION is unique and can be certified without any IPM code: take the active set the oracle's
answer suggests, solve the equality-constrained QP of that active set with one dense numpy KKT solve, and check
(i) the two primal solutions coincide, (ii) every inactive constraint holds, (iii) every active multiplier has the right
sign.  (i)-(iii) are the KKT conditions of the original QP, which are sufficient for optimality of a convex QP.
Problem data conventions follow SURVEY 8a A5/A7 (delta form, t-vector order, slack stationarity)."""
import numpy as np
import pytest

from ad_mpc_b200 import workload as wl
from oracle import oracle as orc

NC = 10


def _build_qp(o, lin, it_x, it_u, x0):
    """Dense QP in v = [du (2N) | dx_1..N (7N) | sl (2N) | su (2N)]:  min 1/2 v'Hv + g'v,  Aeq v = beq,  C v >= d."""
    N, Ts = o.N, o.dt
    nu, nx = 2 * N, 7 * N
    nv = nu + nx + 4 * N
    iu = lambda k: slice(2 * k, 2 * k + 2)
    ix = lambda k: slice(nu + 7 * (k - 1), nu + 7 * k)            # k = 1..N
    isl = lambda k: slice(nu + nx + 2 * k, nu + nx + 2 * k + 2)
    isu = lambda k: slice(nu + nx + 2 * N + 2 * k, nu + nx + 2 * N + 2 * k + 2)
    W, We = np.array(o.W[:]), np.array(o.We[:])
    H, g = np.zeros((nv, nv)), np.zeros(nv)
    d0 = x0 - it_x[0]
    for k in range(N):
        H[iu(k), iu(k)] = np.diag(Ts * W[7:9]); g[iu(k)] = lin["r"][k]
        H[isl(k), isl(k)] = np.diag(Ts * np.array(o.Zl[:])); g[isl(k)] = Ts * np.array(o.zl[:])
        H[isu(k), isu(k)] = np.diag(Ts * np.array(o.Zu[:])); g[isu(k)] = Ts * np.array(o.zu[:])
    for k in range(1, N + 1):
        H[ix(k), ix(k)] = np.diag(Ts * W[:7]) if k < N else np.diag(We)
        g[ix(k)] = lin["q"][k]
    Aeq, beq = np.zeros((nx, nv)), np.zeros(nx)
    for k in range(N):                                             # dx_{k+1} - A_k dx_k - B_k du_k = b_k
        rows = slice(7 * k, 7 * k + 7)
        Aeq[rows, ix(k + 1)] = np.eye(7)
        Aeq[rows, iu(k)] = -lin["B"][k]
        beq[rows] = lin["b"][k]
        if k == 0:
            beq[rows] += lin["A"][0] @ d0
        else:
            Aeq[rows, ix(k)] = -lin["A"][k]
    C, d, tag = [], [], []
    lbu, ubu = np.array(o.lbu[:]), np.array(o.ubu[:])
    for k in range(N):
        for j in range(2):
            row = np.zeros(nv); row[2 * k + j] = 1.0; row[isl(k).start + j] = 1.0
            C.append(row); d.append(lbu[j] - it_u[k, j]); tag.append((k, j))                 # t_lbu = du - lo + sl
            row = np.zeros(nv); row[2 * k + j] = -1.0; row[isu(k).start + j] = 1.0
            C.append(row); d.append(-(ubu[j] - it_u[k, j])); tag.append((k, 3 + j))          # t_ubu = hi - du + su
            row = np.zeros(nv); row[isl(k).start + j] = 1.0
            C.append(row); d.append(0.0); tag.append((k, 6 + j))                             # t_ls = sl
            row = np.zeros(nv); row[isu(k).start + j] = 1.0
            C.append(row); d.append(0.0); tag.append((k, 8 + j))                             # t_us = su
        if k >= 1:
            row = np.zeros(nv); row[ix(k).start + 6] = 1.0
            C.append(row); d.append(o.lbx - it_x[k, 6]); tag.append((k, 2))                  # t_lbx = ddelta - lox
            row = np.zeros(nv); row[ix(k).start + 6] = -1.0
            C.append(row); d.append(-(o.ubx - it_x[k, 6])); tag.append((k, 5))               # t_ubx = hix - ddelta
    return H, g, Aeq, beq, np.array(C), np.array(d), tag, (iu, ix, isl, isu)


@pytest.mark.parametrize("N,p,perturb,seed,steer", [(6, 0.0, 6.0, 1, 0.52), (10, 1.0, 4.0, 2, 0.52), (20, 0.5, 8.0, 3, 0.52),
                                                         (8, 1.0, 0.5, 4, 0.52), (12, 0.0, 6.0, 5, 0.12)])
def test_oracle_qp_solution_is_the_kkt_point_of_its_active_set(N, p, perturb, seed, steer):
    B = 6
    batch = wl.make_batch(B, N, seed=700 + seed, p=p, perturb=perturb)
    rng = np.random.default_rng(seed)
    batch["u_init"] = rng.normal(size=(B, N, 2)) * np.array([2.0, 0.8])            # close to / beyond the input bounds
    batch["x_init"][:, :, 6] += rng.normal(size=(B, N + 1)) * 0.25                  # steering near its hard bound
    o = orc.default_opts(N, lbu=[-3.0, -1.0], ubu=[2.0, 1.0], Zl=[0.0, 0.3], Zu=[0.0, 0.3], lbx=-steer, ubx=steer)
    if steer < 0.5:
        batch["x_init"][:, :, 6] = np.clip(batch["x_init"][:, :, 6], -0.1, 0.1)       # iterate inside the tight steering box
    n_active_total, n_hard = 0, 0
    for b in range(B):
        it = orc.make_iterate(o, batch["x_init"][b], batch["u_init"][b])
        lin = orc.prepare(o, it, batch["yref"][b], batch["p"][b])
        sol = orc.qp_solve(o, lin["_c"], it, batch["x0"][b])
        assert sol["qp_status"] == 0
        H, g, Aeq, beq, C, d, tag, (iu, ix, isl, isu) = _build_qp(o, lin, batch["x_init"][b], batch["u_init"][b], batch["x0"][b])
        v = np.concatenate([sol["du"].reshape(-1), sol["dx"][1:].reshape(-1), sol["sl"].reshape(-1), sol["su"].reshape(-1)])
        # oracle-side quantities in the same order as the rows of C
        t_or = np.array([sol["t"][k, c] for k, c in tag])
        lam_or = np.array([sol["lam"][k, c] for k, c in tag])
        assert np.abs(C @ v - d - t_or).max() < 1e-7                      # t-vector definition
        active = lam_or > t_or                                             # complementarity: lam * t ~ 1e-9 at the solution
        n_active_total += int(active.sum())
        n_hard += sum(1 for (k, c), a in zip(tag, active) if a and c in (2, 5))
        Ca, da = C[active], d[active]
        na, ne, nv = Ca.shape[0], Aeq.shape[0], H.shape[0]
        K = np.zeros((nv + ne + na, nv + ne + na))
        K[:nv, :nv] = H
        K[:nv, nv:nv + ne] = Aeq.T; K[nv:nv + ne, :nv] = Aeq
        K[:nv, nv + ne:] = -Ca.T; K[nv + ne:, :nv] = Ca
        rhs = np.concatenate([-g, beq, da])
        z = np.linalg.lstsq(K, rhs, rcond=None)[0]
        v_as, lam_as = z[:nv], z[nv + ne:]
        assert np.abs(K @ z - rhs).max() < 1e-8                                # the active-set KKT system is consistent
        scale = max(1.0, np.abs(v_as).max())
        assert np.abs(v - v_as).max() <= 2e-7 * scale, np.abs(v - v_as).max()  # (i) same primal point
        assert (C[~active] @ v_as - d[~active]).min() > -1e-7                  # (ii) inactive constraints hold
        assert lam_as.min() > -1e-7                                            # (iii) multipliers of the active set >= 0
        assert np.abs(lam_as - lam_or[active]).max() <= 1e-5 * max(1.0, np.abs(lam_as).max())
    assert n_active_total > 0                                                  # the scenarios do hit their bounds
    if steer < 0.5:
        assert n_hard > 0                                                      # ... including the hard steering bound


def _build_qp_set1(o, lin, it_x, it_u, x0):
    """The same dense QP for the Frenet variant's constraint set (con_set = 1): u0 soft (slack 0), u1 hard, e_y = x[1] hard,
    delta = x[6] soft (slack 1); rows [lb(4) | ub(4) | ls(2) | us(2)], state rows only for k >= 1."""
    N, Ts = o.N, o.dt
    nu, nx = 2 * N, 7 * N
    nv = nu + nx + 4 * N
    iu = lambda k: slice(2 * k, 2 * k + 2)
    ix = lambda k: slice(nu + 7 * (k - 1), nu + 7 * k)
    isl = lambda k: slice(nu + nx + 2 * k, nu + nx + 2 * k + 2)
    isu = lambda k: slice(nu + nx + 2 * N + 2 * k, nu + nx + 2 * N + 2 * k + 2)
    W, We = np.array(o.W[:]), np.array(o.We[:])
    H, g = np.zeros((nv, nv)), np.zeros(nv)
    d0 = x0 - it_x[0]
    for k in range(N):
        H[iu(k), iu(k)] = np.diag(Ts * W[7:9]); g[iu(k)] = lin["r"][k]
        H[isl(k), isl(k)] = np.diag(Ts * np.array(o.Zl[:])); g[isl(k)] = Ts * np.array(o.zl[:])
        H[isu(k), isu(k)] = np.diag(Ts * np.array(o.Zu[:])); g[isu(k)] = Ts * np.array(o.zu[:])
        if k == 0:                                                 # the steering-angle slack does not exist at stage 0: pin it
            H[isl(0).start + 1, isl(0).start + 1] += 1.0; H[isu(0).start + 1, isu(0).start + 1] += 1.0
            g[isl(0).start + 1] = 0.0; g[isu(0).start + 1] = 0.0
    for k in range(1, N + 1):
        H[ix(k), ix(k)] = np.diag(Ts * W[:7]) if k < N else np.diag(We)
        g[ix(k)] = lin["q"][k]
    Aeq, beq = np.zeros((nx, nv)), np.zeros(nx)
    for k in range(N):
        rows = slice(7 * k, 7 * k + 7)
        Aeq[rows, ix(k + 1)] = np.eye(7)
        Aeq[rows, iu(k)] = -lin["B"][k]
        beq[rows] = lin["b"][k]
        if k == 0:
            beq[rows] += lin["A"][0] @ d0
        else:
            Aeq[rows, ix(k)] = -lin["A"][k]
    quant = [("u", 0, 0, o.lbu[0], o.ubu[0]), ("u", 1, None, o.lbu[1], o.ubu[1]), ("x", 1, None, o.lbx2, o.ubx2),
             ("x", 6, 1, o.lbx, o.ubx)]
    C, d, tag = [], [], []
    for k in range(N):
        for q, (kind, idx, sq, lo, hi) in enumerate(quant):
            if kind == "x" and k == 0:
                continue
            col = 2 * k + idx if kind == "u" else ix(k).start + idx
            bar = it_u[k, idx] if kind == "u" else it_x[k, idx]
            row = np.zeros(nv); row[col] = 1.0
            if sq is not None: row[isl(k).start + sq] = 1.0
            C.append(row); d.append(lo - bar); tag.append((k, q))
            row = np.zeros(nv); row[col] = -1.0
            if sq is not None: row[isu(k).start + sq] = 1.0
            C.append(row); d.append(-(hi - bar)); tag.append((k, 4 + q))
            if sq is not None:
                row = np.zeros(nv); row[isl(k).start + sq] = 1.0
                C.append(row); d.append(0.0); tag.append((k, 8 + sq))
                row = np.zeros(nv); row[isu(k).start + sq] = 1.0
                C.append(row); d.append(0.0); tag.append((k, 10 + sq))
    return H, g, Aeq, beq, np.array(C), np.array(d), tag


@pytest.mark.parametrize("N,seed", [(8, 1), (14, 2), (20, 3)])
def test_oracle_qp_solution_with_the_frenet_constraint_set_is_a_kkt_point(N, seed):
    """Same certificate for con_set = 1 (u0 soft, u1 hard, e_y hard, steering angle soft): the generic-constraint IPM's answer
    is the KKT point of its active set, with hard steering-rate / e_y rows and soft steering-angle rows active."""
    B = 6
    batch = wl.make_batch_frenet(B, N, seed=900 + seed, p=1.0, perturb=3.0)
    rng = np.random.default_rng(seed)
    batch["u_init"] = rng.normal(size=(B, N, 2)) * np.array([1.0, 0.2])
    batch["x_init"][:, :, 1] = np.clip(batch["x_init"][:, :, 1] + rng.normal(size=(B, N + 1)) * 0.3, -0.5, 0.5)
    batch["x_init"][:, :, 6] += rng.normal(size=(B, N + 1)) * 0.05
    batch["x0"][:, 1] = np.clip(batch["x0"][:, 1], -0.5, 0.5)
    q = [0.0, 10.0, 10.0, 10.0, 10.0, 1.0, 0.1]
    o = orc.default_opts(N, con_set=1, model_backend=2, W=q + [10.0, 10.0], We=[0.01 * v for v in q], zl=[100.0, 100.0],
                         zu=[100.0, 100.0], Zl=[0.0, 0.2], Zu=[0.0, 0.2], lbu=[-2.0, -0.5], ubu=[1.5, 0.5], lbx=-0.08, ubx=0.08,
                         lbx2=-0.6, ubx2=0.6)
    n_cert, n_hard_u1, n_hard_ey, n_soft_delta = 0, 0, 0, 0
    for b in range(B):
        it = orc.make_iterate(o, batch["x_init"][b], batch["u_init"][b])
        lin = orc.prepare(o, it, batch["yref"][b], batch["p"][b])          # (Cartesian linearisation: any QP data will do)
        sol = orc.qp_solve(o, lin["_c"], it, batch["x0"][b])
        if sol["qp_status"] != 0 or max(np.abs(sol["du"]).max(), np.abs(sol["dx"]).max()) > 50.0:
            continue                    # infeasible hard boxes: nothing to certify ; runaway soft acceleration: ill-conditioned QP
        H, g, Aeq, beq, C, d, tag = _build_qp_set1(o, lin, batch["x_init"][b], batch["u_init"][b], batch["x0"][b])
        v = np.concatenate([sol["du"].reshape(-1), sol["dx"][1:].reshape(-1), sol["sl"].reshape(-1), sol["su"].reshape(-1)])
        t_or = np.array([sol["t"][k, c] for k, c in tag])
        lam_or = np.array([sol["lam"][k, c] for k, c in tag])
        assert np.abs(C @ v - d - t_or).max() < 1e-7
        active = lam_or > t_or
        n_hard_u1 += sum(1 for (k, c), a in zip(tag, active) if a and c in (1, 5))
        n_hard_ey += sum(1 for (k, c), a in zip(tag, active) if a and c in (2, 6))
        n_soft_delta += sum(1 for (k, c), a in zip(tag, active) if a and c in (3, 7))
        Ca, da = C[active], d[active]
        na, ne, nv = Ca.shape[0], Aeq.shape[0], H.shape[0]
        K = np.zeros((nv + ne + na, nv + ne + na))
        K[:nv, :nv] = H
        K[:nv, nv:nv + ne] = Aeq.T; K[nv:nv + ne, :nv] = Aeq
        K[:nv, nv + ne:] = -Ca.T; K[nv + ne:, :nv] = Ca
        rhs = np.concatenate([-g, beq, da])
        z = np.linalg.lstsq(K, rhs, rcond=None)[0]
        v_as, lam_as = z[:nv], z[nv + ne:]
        assert np.abs(K @ z - rhs).max() < 1e-9 * max(1.0, np.abs(z).max())     # consistent (soft acceleration: large steps)
        scale = max(1.0, np.abs(v_as).max())
        assert np.abs(v - v_as).max() <= 2e-7 * scale, np.abs(v - v_as).max()
        assert (C[~active] @ v_as - d[~active]).min() > -1e-7
        assert lam_as.min() > -1e-7
        assert np.abs(lam_as - lam_or[active]).max() <= 1e-5 * max(1.0, np.abs(lam_as).max())
        n_cert += 1
    assert n_cert >= 3 and n_hard_u1 > 0 and n_soft_delta > 0
