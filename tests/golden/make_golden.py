"""Generates the committed golden fixtures from the reference tree (run in the build container only).

  python tests/golden/make_golden.py

Inputs (read-only, never copied as source):
  /root/reference/data_driven_mpc/ros_gp_mpc/src/ad_mpc/sim_car_iterate.json   acados store_iterate dump (N=40, p=0,
      converged) -- the only artefact in the reference that pins acados/HPIPM output for this path
  /root/reference/.../solve_iteration.json                                     layout fixture (non-converged, p~1)
  oracle/_ref/libsim_car_ref.so    the reference's CasADi-generated sim_car_expl_vde_forw / _ode_fun, compiled unmodified

Outputs:
  tests/golden/sim_car_iterate.npz   x[41,7] u[40,2] pi[40,7] lam0[22] t0[22] lam[39,10] t[39,10] sl[40,2] su[40,2]
  tests/golden/solve_iteration.npz   same keys
  tests/golden/vde_vectors.npz       300 seeded (x,Sx,Su,u,p) -> (xdot,dSx,dSu) evaluations of the reference VDE/ODE
"""
import ctypes as C
import json
import os

import numpy as np

REF = "/root/reference/data_driven_mpc/ros_gp_mpc/src/ad_mpc"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def convert_iterate(name):
    d = json.load(open(os.path.join(REF, name + ".json")))
    N = 40
    out = dict(
        x=np.array([d["x_%d" % k] for k in range(N + 1)]),
        u=np.array([d["u_%d" % k] for k in range(N)]),
        pi=np.array([d["pi_%d" % k] for k in range(N)]),
        lam0=np.array(d["lam_0"]), t0=np.array(d["t_0"]),
        lam=np.array([d["lam_%d" % k] for k in range(1, N)]),
        t=np.array([d["t_%d" % k] for k in range(1, N)]),
        sl=np.array([d["sl_%d" % k] for k in range(N)]),
        su=np.array([d["su_%d" % k] for k in range(N)]),
    )
    assert d["lam_40"] == [] and d["u_40"] == []
    np.savez(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def vde_vectors():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsim_car_ref.so"))
    rng = np.random.default_rng(20260)
    n = 300
    X = rng.normal(size=(n, 7)) * np.array([20, 20, 1.5, 1, 0.5, 0.3, 0.25])
    X[:, 3] = rng.uniform(0.5, 15, size=n)
    U = rng.normal(size=(n, 2)) * np.array([3, 1])
    P = rng.choice([0.0, 1.0, 0.25, 0.7], size=n)
    SX = rng.normal(size=(n, 7, 7))
    SU = rng.normal(size=(n, 7, 2))
    SX[:100] = np.eye(7)
    SU[:100] = 0
    xdot, dsx, dsu, ode = np.zeros((n, 7)), np.zeros((n, 7, 7)), np.zeros((n, 7, 2)), np.zeros((n, 7))
    dp = C.POINTER(C.c_double)
    for i in range(n):
        sxc = np.asfortranarray(SX[i]).ravel(order="F").copy()
        suc = np.asfortranarray(SU[i]).ravel(order="F").copy()
        p = np.array([P[i]])
        o0, o1, o2 = np.zeros(7), np.zeros(42), np.zeros(13)
        arg = (dp * 12)(*[a.ctypes.data_as(dp) for a in (X[i], sxc, suc, U[i], p)])
        res = (dp * 10)(*[a.ctypes.data_as(dp) for a in (o0, o1, o2)])
        iw = (C.c_int * 3)()
        w = (C.c_double * 244)()
        lib.sim_car_expl_vde_forw(arg, res, iw, w, 0)
        xdot[i] = o0
        dsx[i, :6, :] = o1.reshape(7, 6).T       # CCS: 6 rows per column (sim_car_expl_vde_forw.c:118)
        dsu[i, :6, 0] = o2[:6]
        dsu[i, :, 1] = o2[6:]
        arg = (dp * 12)(*[a.ctypes.data_as(dp) for a in (X[i], U[i], p)])
        res = (dp * 10)(ode[i].ctypes.data_as(dp))
        lib.sim_car_expl_ode_fun(arg, res, iw, w, 0)
    np.savez(os.path.join(HERE, "vde_vectors.npz"), x=X, u=U, p=P, Sx=SX, Su=SU, xdot=xdot, dSx=dsx, dSu=dsu, ode=ode)
    print("vde vectors", n, "max|ode-xdot|", np.abs(ode - xdot).max())


if __name__ == "__main__":
    convert_iterate("sim_car_iterate")
    convert_iterate("solve_iteration")
    vde_vectors()
