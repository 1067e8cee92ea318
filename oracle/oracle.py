"""ctypes wrapper around oracle/liboracle.so -- CPU ORACLE, test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product (ad_mpc_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NX, NU, NC, NMAX, DZMAX, GPOUT_MAX = 7, 2, 10, 128, 8, 4
NCS = 12      # storage rows per stage (ORC_NC); rows in use / stride = con_rows(o): 10 Cartesian set, 12 Frenet set


class OrcOpts(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("iter_max", C.c_int), ("gp_enabled", C.c_int), ("gp_nout", C.c_int),
        ("gp_M", C.c_int), ("gp_dz", C.c_int), ("gp_stage0_trigger", C.c_int), ("model_backend", C.c_int),
        ("gp_feat", C.c_int * DZMAX), ("gp_row", C.c_int * GPOUT_MAX),
        ("dt", C.c_double), ("W", C.c_double * 9), ("We", C.c_double * 7),
        ("zl", C.c_double * 2), ("zu", C.c_double * 2), ("Zl", C.c_double * 2), ("Zu", C.c_double * 2),
        ("lbu", C.c_double * 2), ("ubu", C.c_double * 2), ("lbx", C.c_double), ("ubx", C.c_double),
        ("mass", C.c_double), ("lf", C.c_double), ("lr", C.c_double), ("iz", C.c_double),
        ("cf2", C.c_double), ("cr2", C.c_double),
        ("mu0", C.c_double), ("tol_stat", C.c_double), ("tol_eq", C.c_double), ("tol_ineq", C.c_double),
        ("tol_comp", C.c_double), ("alpha_min", C.c_double), ("lam_min", C.c_double), ("t_min", C.c_double),
        ("thr0", C.c_double), ("reg", C.c_double),
        ("con_set", C.c_int), ("lbx2", C.c_double), ("ubx2", C.c_double),
    ]


class OrcGp(C.Structure):
    _fields_ = [("X", C.c_void_p), ("alpha", C.c_void_p), ("ell", C.c_void_p), ("sigma_f", C.c_void_p),
                ("y_mean", C.c_void_p)]


class OrcIterate(C.Structure):
    _fields_ = [("x", C.c_double * ((NMAX + 1) * NX)), ("u", C.c_double * (NMAX * NU)),
                ("pi", C.c_double * (NMAX * NX)), ("lam", C.c_double * (NMAX * NCS)),
                ("t", C.c_double * (NMAX * NCS)), ("sl", C.c_double * (NMAX * NU)), ("su", C.c_double * (NMAX * NU))]


class OrcLin(C.Structure):
    _fields_ = [("A", C.c_double * (NMAX * 49)), ("B", C.c_double * (NMAX * 14)), ("b", C.c_double * (NMAX * 7)),
                ("q", C.c_double * ((NMAX + 1) * 7)), ("r", C.c_double * (NMAX * 2))]


class OrcStats(C.Structure):
    _fields_ = [("status", C.c_int), ("qp_status", C.c_int), ("qp_iter", C.c_int), ("res", C.c_double * 4),
                ("step_inf", C.c_double)]


class OrcQpSol(C.Structure):
    _fields_ = [("dx", C.c_double * ((NMAX + 1) * 7)), ("du", C.c_double * (NMAX * 2)), ("pi", C.c_double * (NMAX * 7)),
                ("lam", C.c_double * (NMAX * NCS)), ("t", C.c_double * (NMAX * NCS)),
                ("sl", C.c_double * (NMAX * 2)), ("su", C.c_double * (NMAX * 2))]


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present). Building the checker is not using it."""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("rti_oracle.c", "rti_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(os.path.join(_HERE, "_ref", "libsim_car_ref.so"))):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


_lib = None
_so_override = None


def use_native():
    """bench.py's CPU legs only: rebuild the restatement with -O3 -march=native ON THE MACHINE THAT RUNS IT (the shipped
    liboracle.so is built -march=x86-64-v3 in the build container so that it loads on any box).  The result lives in
    oracle/_native/ (git- and gpurun-ignored: it must never travel).  Returns the flags in use."""
    global _lib, _so_override
    d = os.path.join(_HERE, "_native")
    so = os.path.join(d, "liboracle_native.so")
    flags = "-O3 -march=native"
    try:
        os.makedirs(d, exist_ok=True)
        subprocess.check_call(["/usr/bin/gcc"] + flags.split() + ["-fPIC", "-fopenmp", "-std=c11", "-shared", "-o", so,
                               os.path.join(_HERE, "rti_oracle.c"), "-lm", "-ldl"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    except Exception:
        return "-O3 -march=x86-64-v3 (native rebuild failed; shipped build)"
    _so_override, _lib = so, None
    return flags


def lib():
    global _lib
    if _lib is None:
        so = _so_override or build()
        L = C.CDLL(so)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_default_opts.argtypes = [C.POINTER(OrcOpts)]
        L.orc_con_rows.argtypes = [C.POINTER(OrcOpts)]
        L.orc_con_check.argtypes = [C.POINTER(OrcOpts), C.POINTER(OrcIterate), dp]
        L.orc_model_jac.argtypes = [C.POINTER(OrcOpts), C.POINTER(OrcGp), dp, dp, C.c_double, dp, C.c_double, dp, dp, dp]
        L.orc_gp_predict.argtypes = [C.POINTER(OrcOpts), C.POINTER(OrcGp), dp, dp, dp]
        L.orc_rk4_sens.argtypes = [C.POINTER(OrcOpts), C.POINTER(OrcGp), dp, dp, C.c_double, dp, C.c_double, dp, dp, dp]
        L.orc_rk4_sens.restype = C.c_int
        L.orc_prepare.argtypes = [C.POINTER(OrcOpts), C.POINTER(OrcGp), C.POINTER(OrcIterate), dp, dp, dp, C.POINTER(OrcLin)]
        L.orc_qp_solve.argtypes = [C.POINTER(OrcOpts), C.POINTER(OrcLin), C.POINTER(OrcIterate), dp, C.POINTER(OrcQpSol), C.POINTER(OrcStats)]
        L.orc_rti_step.argtypes = [C.POINTER(OrcOpts), C.POINTER(OrcGp), dp, dp, dp, dp, C.POINTER(OrcIterate), C.POINTER(OrcStats)]
        L.orc_rti_batch.argtypes = [C.POINTER(OrcOpts), C.POINTER(OrcGp), C.c_int, dp, dp, dp, dp, dp, dp, dp, ip, ip, ip, C.c_int]
        L.orc_set_kappa.argtypes = [C.c_double]
        L.orc_set_kappa_spline.argtypes = [C.c_int, dp, dp]
        L.orc_set_batch_kappa_spline.argtypes = [C.c_int, dp, dp]
        L.orc_rti_batch_frenet.argtypes = [C.POINTER(OrcOpts), C.POINTER(OrcGp), C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, ip, ip,
                                           ip, C.c_int]
        L.orc_sqp_batch_frenet.argtypes = [C.POINTER(OrcOpts), C.POINTER(OrcGp), C.c_int, dp, dp, dp, dp, dp, dp, dp, C.c_int, dp,
                                           ip, ip, dp, C.c_int]
        L.orc_load_ref_model.argtypes = [C.c_char_p]
        _lib = L
    return _lib


def ref_model_path():
    return os.path.join(_HERE, "_ref", "libsim_car_ref.so")


def load_ref_model():
    """dlopen the compiled reference CasADi model (oracle/_ref). Returns True when available."""
    p = ref_model_path()
    return os.path.exists(p) and lib().orc_load_ref_model(p.encode()) == 0


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def con_rows(o):
    """Inequality rows per stage of the configured constraint set (stride of lam / t)."""
    return lib().orc_con_rows(C.byref(o))


def default_opts(N=20, **kw):
    o = OrcOpts()
    lib().orc_default_opts(C.byref(o))
    o.N = N
    for k, v in kw.items():
        cur = getattr(o, k)
        if hasattr(cur, "__len__"):
            for i, vi in enumerate(v):
                cur[i] = vi
        else:
            setattr(o, k, v)
    return o


class Gp:
    """Holds GP arrays alive and exposes the C struct. model: dict with X[nout,M,dz], alpha[nout,M], ell[nout,dz],
    sigma_f[nout], y_mean[nout] (float64)."""

    def __init__(self, model):
        self.X = np.ascontiguousarray(model["X"], dtype=np.float64)
        self.alpha = np.ascontiguousarray(model["alpha"], dtype=np.float64)
        self.ell = np.ascontiguousarray(model["ell"], dtype=np.float64)
        self.sigma_f = np.ascontiguousarray(model["sigma_f"], dtype=np.float64)
        self.y_mean = np.ascontiguousarray(model["y_mean"], dtype=np.float64)
        self.c = OrcGp(self.X.ctypes.data, self.alpha.ctypes.data, self.ell.ctypes.data,
                       self.sigma_f.ctypes.data, self.y_mean.ctypes.data)

    def apply(self, o, feat=(3, 4, 5, 6), rows=(4, 5), stage0_trigger=1):
        o.gp_enabled = 1
        o.gp_nout, o.gp_M, o.gp_dz = self.X.shape
        for i, f in enumerate(feat):
            o.gp_feat[i] = f
        for i, r in enumerate(rows):
            o.gp_row[i] = r
        o.gp_stage0_trigger = stage0_trigger
        return o


_spline_keep = []


def set_kappa_spline(breaks=None, coef=None):
    """Direct calls (model_jac / rk4_sens on this thread): kappa(s) = piecewise cubic, breaks[K+1], coef[K,4] (lowest power
    first); None switches back to the per-node constant."""
    global _spline_keep
    if breaks is None:
        lib().orc_set_kappa_spline(0, None, None)
        _spline_keep = []
        return
    b = np.ascontiguousarray(breaks, dtype=np.float64); c = np.ascontiguousarray(coef, dtype=np.float64)
    _spline_keep = [b, c]
    lib().orc_set_kappa_spline(c.shape[0], _dp(b), _dp(c))


def set_batch_kappa_spline(breaks=None, coef=None):
    """Batch entry points (rti_batch / sqp_batch with the Frenet backend): per-instance splines breaks[B,K+1], coef[B,K,4]."""
    global _spline_keep
    if breaks is None:
        lib().orc_set_batch_kappa_spline(0, None, None)
        return
    b = np.ascontiguousarray(breaks, dtype=np.float64); c = np.ascontiguousarray(coef, dtype=np.float64)
    _spline_keep = [b, c]
    lib().orc_set_batch_kappa_spline(c.shape[1], _dp(b), _dp(c))


def model_jac(o, x, u, p, gp=None, gp_state=None, trigger=0.0, kappa=0.0):
    """kappa: path curvature, used by the Frenet variant (o.model_backend == 2) only."""
    f, Jx, Ju = np.zeros(7), np.zeros((7, 7)), np.zeros((7, 2))
    lib().orc_set_kappa(float(kappa))
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    gs = None if gp_state is None else np.ascontiguousarray(gp_state, dtype=np.float64)
    lib().orc_model_jac(C.byref(o), C.byref(gp.c) if gp else None, _dp(x), _dp(u), float(p), _dp(gs), float(trigger),
                        _dp(f), _dp(Jx), _dp(Ju))
    return f, Jx, Ju


def gp_predict(o, gp, z):
    z = np.ascontiguousarray(z, dtype=np.float64)
    mu, dmu = np.zeros(o.gp_nout), np.zeros((o.gp_nout, o.gp_dz))
    lib().orc_gp_predict(C.byref(o), C.byref(gp.c), _dp(z), _dp(mu), _dp(dmu))
    return mu, dmu


def rk4_sens(o, x, u, p, gp=None, gp_state=None, trigger=0.0, kappa=0.0):
    xn, A, B = np.zeros(7), np.zeros((7, 7)), np.zeros((7, 2))
    lib().orc_set_kappa(float(kappa))
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    gs = None if gp_state is None else np.ascontiguousarray(gp_state, dtype=np.float64)
    bad = lib().orc_rk4_sens(C.byref(o), C.byref(gp.c) if gp else None, _dp(x), _dp(u), float(p), _dp(gs),
                             float(trigger), _dp(xn), _dp(A), _dp(B))
    return xn, A, B, bad


def make_iterate(o, x=None, u=None):
    it = OrcIterate()
    N = o.N
    if x is not None:
        it.x[:(N + 1) * 7] = list(np.asarray(x, dtype=np.float64).reshape(-1))
    if u is not None:
        it.u[:N * 2] = list(np.asarray(u, dtype=np.float64).reshape(-1))
    return it


def set_iterate_duals(o, it, lam=None, t=None, sl=None, su=None, pi=None):
    """Fill the dual part of an iterate: lam, t [N, con_rows(o)], sl, su [N, 2], pi [N, 7]."""
    N = o.N
    for name, a, w in (("lam", lam, con_rows(o)), ("t", t, con_rows(o)), ("sl", sl, 2), ("su", su, 2), ("pi", pi, 7)):
        if a is not None:
            getattr(it, name)[:N * w] = list(np.asarray(a, dtype=np.float64).reshape(N * w))
    return it


def con_check(o, it):
    """(max |t - constraint function|, max |slack stationarity|, max |lam t|) of an iterate, model-independent."""
    out = np.zeros(3)
    lib().orc_con_check(C.byref(o), C.byref(it), _dp(out))
    return out


def iterate_arrays(o, it):
    N = o.N
    g = lambda f, n, w: np.array(f[:n * w]).reshape(n, w)
    return dict(x=g(it.x, N + 1, 7), u=g(it.u, N, 2), pi=g(it.pi, N, 7), lam=g(it.lam, N, con_rows(o)), t=g(it.t, N, con_rows(o)),
                sl=g(it.sl, N, 2), su=g(it.su, N, 2))


def prepare(o, it, yref, p, gp=None, gp_state=None):
    lin = OrcLin()
    yref = np.ascontiguousarray(yref, dtype=np.float64).reshape(-1)
    p = np.ascontiguousarray(p, dtype=np.float64)
    gs = None if gp_state is None else np.ascontiguousarray(gp_state, dtype=np.float64)
    bad = lib().orc_prepare(C.byref(o), C.byref(gp.c) if gp else None, C.byref(it), _dp(yref), _dp(p), _dp(gs), C.byref(lin))
    N = o.N
    return dict(A=np.array(lin.A[:N * 49]).reshape(N, 7, 7), B=np.array(lin.B[:N * 14]).reshape(N, 7, 2),
                b=np.array(lin.b[:N * 7]).reshape(N, 7), q=np.array(lin.q[:(N + 1) * 7]).reshape(N + 1, 7),
                r=np.array(lin.r[:N * 2]).reshape(N, 2), bad=bad, _c=lin)


def qp_solve(o, lin_c, it, x0):
    sol, st = OrcQpSol(), OrcStats()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    lib().orc_qp_solve(C.byref(o), C.byref(lin_c), C.byref(it), _dp(x0), C.byref(sol), C.byref(st))
    N = o.N
    g = lambda f, n, w: np.array(f[:n * w]).reshape(n, w)
    return dict(dx=g(sol.dx, N + 1, 7), du=g(sol.du, N, 2), pi=g(sol.pi, N, 7), lam=g(sol.lam, N, con_rows(o)),
                t=g(sol.t, N, con_rows(o)), sl=g(sol.sl, N, 2), su=g(sol.su, N, 2), qp_status=st.qp_status, qp_iter=st.qp_iter,
                res=np.array(st.res[:]))


def rti_step(o, it, x0, yref, p, gp=None, gp_state=None):
    st = OrcStats()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    yref = np.ascontiguousarray(yref, dtype=np.float64).reshape(-1)
    p = np.ascontiguousarray(p, dtype=np.float64)
    gs = None if gp_state is None else np.ascontiguousarray(gp_state, dtype=np.float64)
    lib().orc_rti_step(C.byref(o), C.byref(gp.c) if gp else None, _dp(x0), _dp(yref), _dp(p), _dp(gs), C.byref(it), C.byref(st))
    return dict(status=st.status, qp_status=st.qp_status, qp_iter=st.qp_iter, res=np.array(st.res[:]), step_inf=st.step_inf)


def rti_batch(o, x0, yref, p, xit, uit, gp=None, gp_state=None, nthreads=0, kappa=None):
    """x0[B,7] yref[B,N*9+7] p[B,N] xit[B,N+1,7] uit[B,N,2] -> dict(x,u,pi,status,qp_status,qp_iter); inputs not modified.
    kappa[B,N] (Frenet variant, o.model_backend == 2): path curvature at every shooting node."""
    B = x0.shape[0]
    N = o.N
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    yref = np.ascontiguousarray(yref, dtype=np.float64).reshape(B, -1)
    p = np.ascontiguousarray(np.broadcast_to(np.asarray(p, dtype=np.float64).reshape(B, -1), (B, N)))
    x = np.array(xit, dtype=np.float64, order="C").reshape(B, N + 1, 7)
    u = np.array(uit, dtype=np.float64, order="C").reshape(B, N, 2)
    pi = np.zeros((B, N, 7))
    gs = None if gp_state is None else np.ascontiguousarray(gp_state, dtype=np.float64)
    status, qps, qpi = (np.zeros(B, dtype=np.int32) for _ in range(3))
    kap = None if kappa is None else np.ascontiguousarray(np.broadcast_to(np.asarray(kappa, dtype=np.float64).reshape(B, -1), (B, N)))
    lib().orc_rti_batch_frenet(C.byref(o), C.byref(gp.c) if gp else None, B, _dp(x0), _dp(yref), _dp(p), _dp(kap), _dp(gs),
                               _dp(x), _dp(u), _dp(pi), _ip(status), _ip(qps), _ip(qpi), int(nthreads))
    return dict(x=x, u=u, pi=pi, status=status, qp_status=qps, qp_iter=qpi)


def sqp_batch(o, x0, yref, p, xit, uit, gp=None, gp_state=None, max_iter=100, tol=(1e-6, 1e-6, 1e-6, 1e-6), nthreads=0,
              kappa=None):
    """Full SQP (nlp_solver_type "SQP") on a batch -> dict(x,u,status,sqp_iter,res[B,4]); inputs not modified."""
    B = x0.shape[0]
    N = o.N
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    yref = np.ascontiguousarray(yref, dtype=np.float64).reshape(B, -1)
    p = np.ascontiguousarray(np.broadcast_to(np.asarray(p, dtype=np.float64).reshape(B, -1), (B, N)))
    x = np.array(xit, dtype=np.float64, order="C").reshape(B, N + 1, 7)
    u = np.array(uit, dtype=np.float64, order="C").reshape(B, N, 2)
    gs = None if gp_state is None else np.ascontiguousarray(gp_state, dtype=np.float64)
    status, it = (np.zeros(B, dtype=np.int32) for _ in range(2))
    res = np.zeros((B, 4))
    tol = np.ascontiguousarray(tol, dtype=np.float64)
    kap = None if kappa is None else np.ascontiguousarray(np.broadcast_to(np.asarray(kappa, dtype=np.float64).reshape(B, -1), (B, N)))
    lib().orc_sqp_batch_frenet(C.byref(o), C.byref(gp.c) if gp else None, B, _dp(x0), _dp(yref), _dp(p), _dp(kap), _dp(gs),
                               _dp(x), _dp(u), int(max_iter), _dp(tol), _ip(status), _ip(it), _dp(res), int(nthreads))
    return dict(x=x, u=u, status=status, sqp_iter=it, res=res)
