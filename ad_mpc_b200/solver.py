"""Host-side mirror of the reference's solver interface, on top of the C ABI (include/admpc.h).

* ``AcadosOcpSolverB200`` has the methods the reference calls on ``acados_template.AcadosOcpSolver``
  (data_driven_mpc/ros_gp_mpc/src/ad_mpc/ad_3d_optimizer.py:209,430,438,441-442,450,456,462-465):
  ``set(stage, field, value)``, ``solve() -> int``, ``get(stage, field) -> ndarray``, ``get_stats``,
  ``store_iterate/load_iterate`` -- so ``AD3DOptimizer.run_optimization`` runs unchanged against it.
* ``BatchSolver`` is the batched twin: one call solves B instances.

numpy + ctypes only; no torch, no CPU fallback (``_lib.load`` raises when the CUDA library is missing).
"""
import ctypes as C
import json

import numpy as np

from . import _lib
from ._lib import AdmpcOpts, check

NC = 10


def default_opts(N=20, **kw):
    L = _lib.load()
    o = AdmpcOpts()
    L.admpc_default_opts(C.byref(o))
    o.N = N
    for k, v in kw.items():
        cur = getattr(o, k)
        if hasattr(cur, "__len__"):
            for i, vi in enumerate(v):
                cur[i] = vi
        else:
            setattr(o, k, v)
    return o


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def kappa_pp_from_knots(s_knots, curv):
    """Piecewise-cubic form (breaks[K+1], coef[K,4] lowest power first) of the not-a-knot cubic spline through
    (s_knots, curv) -- the degree-3 interpolating spline CasADi's interpolant(..., 'bspline', ...) builds from the same
    data.  Host-side helper for BatchSolver.set_kappa_spline."""
    from scipy.interpolate import CubicSpline
    cs = CubicSpline(np.asarray(s_knots, dtype=np.float64), np.asarray(curv, dtype=np.float64), bc_type="not-a-knot")
    return cs.x.copy(), np.ascontiguousarray(cs.c[::-1].T)


class PinnedArray:
    """float64/int32 numpy view over cudaHostAlloc'ed memory (truly asynchronous H2D/D2H)."""

    def __init__(self, shape, dtype=np.float64):
        L = _lib.load()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._ptr = L.admpc_host_alloc(max(n, 16))
        if not self._ptr:
            raise _lib.AdmpcError("cudaHostAlloc failed: " + L.admpc_last_error().decode())
        buf = (C.c_char * max(n, 16)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._ptr:
            _lib.load().admpc_host_free(self._ptr)
            self._ptr = None


class BatchSolver:
    """B independent SQP-RTI NMPC instances on one GPU."""

    def __init__(self, B, opts=None, device=0, N=None):
        self.L = _lib.load()
        self.opts = opts if opts is not None else default_opts(N or 20)
        if N is not None:
            self.opts.N = N
        self.B, self.N = int(B), int(self.opts.N)
        self.nc = 12 if self.opts.con_set == 1 else NC      # inequality rows per stage (stride of lam / t)
        h = C.c_void_p()
        check(self.L.admpc_batch_create(C.byref(self.opts), self.B, int(device), C.byref(h)), "admpc_batch_create")
        self.h = h
        self._owner = True
        self._keep = []

    @classmethod
    def _view(cls, handle, B, opts):
        """Non-owning wrapper of a chunk handle that lives inside an admpc_pipe."""
        self = cls.__new__(cls)
        self.L = _lib.load()
        self.opts, self.B, self.N = opts, int(B), int(opts.N)
        self.nc = 12 if opts.con_set == 1 else NC
        self.h = C.c_void_p(handle)
        self._owner = False
        self._keep = []
        return self

    def close(self):
        if getattr(self, "h", None):
            if self._owner:
                self.L.admpc_batch_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- model / inputs -------------------------------------------------------------------------------------
    def set_gp(self, model, stage0_trigger=1):
        """model: dict X[nout,M,dz] alpha[nout,M] ell[nout,dz] sigma_f[nout] y_mean[nout] feat rows (or None)."""
        if model is None:
            check(self.L.admpc_batch_set_gp(self.h, 0, 0, 0, None, None, None, None, None, None, None, 0), "set_gp")
            return
        X = _f64(model["X"])
        nout, M, dz = X.shape
        feat = np.ascontiguousarray(model.get("feat", (3, 4, 5, 6)[:dz]), dtype=np.int32)
        rows = np.ascontiguousarray(model.get("rows", (4, 5)[:nout]), dtype=np.int32)
        check(self.L.admpc_batch_set_gp(self.h, nout, M, dz, _ip(feat), _ip(rows), _dp(X), _dp(_f64(model["alpha"])),
                                        _dp(_f64(model["ell"])), _dp(_f64(model["sigma_f"])), _dp(_f64(model["y_mean"])),
                                        int(stage0_trigger)), "admpc_batch_set_gp")

    def set_gp_ensemble(self, models, centroids=None, stage0_trigger=1):
        """GP ensemble (GPEnsemble, gp.py:536-770): `models` = list of K model dicts of the set_gp schema (one per cluster,
        same nout / M / dz / feat / rows), `centroids` [K, dz] cluster means in feature space (default: model["mean"]).
        Clusters are sorted by the first centroid coordinate like GPEnsemble.add_model does (gp.py:592-595)."""
        K = len(models)
        cen = np.asarray([m["mean"] for m in models] if centroids is None else centroids, dtype=np.float64).reshape(K, -1)
        order = np.argsort(cen[:, 0], kind="stable")
        models = [models[i] for i in order]
        cen = _f64(cen[order])
        X = _f64(np.stack([m["X"] for m in models]))
        _, nout, M, dz = X.shape
        feat = np.ascontiguousarray(models[0].get("feat", (3, 4, 5, 6)[:dz]), dtype=np.int32)
        rows = np.ascontiguousarray(models[0].get("rows", (4, 5)[:nout]), dtype=np.int32)
        stack = lambda k: _f64(np.stack([np.asarray(m[k], dtype=np.float64) for m in models]))
        check(self.L.admpc_batch_set_gp_ensemble(self.h, K, nout, M, dz, _ip(feat), _ip(rows), _dp(X), _dp(stack("alpha")),
                                                 _dp(stack("ell")), _dp(stack("sigma_f")), _dp(stack("y_mean")), _dp(cen),
                                                 int(stage0_trigger)), "admpc_batch_set_gp_ensemble")
        return order

    def select_gp(self, x=None, u=None):
        """Nearest-centroid model choice per instance (GPEnsemble.select_gp, gp.py:738-770) from the query state x [B,7]
        (default: the current x0) and input u [B,2] (default: zeros); returns the chosen indices [B]."""
        check(self.L.admpc_batch_select_gp(self.h, None if x is None else _dp(_f64(x, (self.B, 7))),
                                           None if u is None else _dp(_f64(u, (self.B, 2)))), "select_gp")
        return self.get_gp_index()

    def set_gp_index(self, idx):
        """Explicit model choice per instance (the reference's use_model argument)."""
        i = np.ascontiguousarray(np.broadcast_to(np.asarray(idx, dtype=np.int32).reshape(-1), (self.B,)))
        check(self.L.admpc_batch_set_gp_index(self.h, _ip(i)), "set_gp_index")

    def get_gp_index(self):
        out = np.empty(self.B, dtype=np.int32)
        check(self.L.admpc_batch_get_gp_index(self.h, _ip(out)), "get_gp_index")
        return out

    def set_x0(self, x0):
        check(self.L.admpc_batch_set_x0(self.h, _dp(_f64(x0, (self.B, 7)))), "set_x0")

    def set_yref(self, yref):
        check(self.L.admpc_batch_set_yref(self.h, _dp(_f64(yref, (self.B, self.N * 9 + 7)))), "set_yref")

    def set_p(self, p):
        p = np.asarray(p, dtype=np.float64)
        if p.size == self.B:
            check(self.L.admpc_batch_set_p_scalar(self.h, _dp(_f64(p, (self.B,)))), "set_p")
        else:
            check(self.L.admpc_batch_set_p(self.h, _dp(_f64(p, (self.B, self.N)))), "set_p")

    def set_kappa(self, kappa):
        """Frenet variant (default_opts(model_variant=1)): path curvature at every shooting node, [B, N] or [B]."""
        k = np.asarray(kappa, dtype=np.float64)
        if k.size == self.B:
            k = np.repeat(k.reshape(self.B, 1), self.N, axis=1)
        check(self.L.admpc_batch_set_kappa(self.h, _dp(_f64(k, (self.B, self.N)))), "set_kappa")

    def set_kappa_spline(self, breaks=None, coef=None):
        """Frenet variant: kappa(s) as per-instance piecewise cubics, breaks[B,K+1] (or [K+1]), coef[B,K,4] (or [K,4]),
        lowest power first -- evaluated inside the model with its d kappa / d s column (the reference's bspline
        interpolant semantics).  None switches back to the per-node constants.  `kappa_pp_from_knots` builds the arrays."""
        if breaks is None:
            check(self.L.admpc_batch_set_kappa_spline(self.h, 0, None, None), "set_kappa_spline")
            return
        c = np.asarray(coef, dtype=np.float64)
        K = c.shape[-2]
        b = _f64(np.broadcast_to(np.asarray(breaks, dtype=np.float64), (self.B, K + 1)))
        c = _f64(np.broadcast_to(c, (self.B, K, 4)))
        check(self.L.admpc_batch_set_kappa_spline(self.h, K, _dp(b), _dp(c)), "set_kappa_spline")

    def set_gp_state(self, gp_state):
        check(self.L.admpc_batch_set_gp_state(self.h, None if gp_state is None else _dp(_f64(gp_state, (self.B, 7)))), "set_gp_state")

    def set_iterate(self, x=None, u=None):
        check(self.L.admpc_batch_set_iterate(self.h, None if x is None else _dp(_f64(x, (self.B, (self.N + 1) * 7))),
                                             None if u is None else _dp(_f64(u, (self.B, self.N * 2)))), "set_iterate")

    def set_duals(self, pi=None, lam=None, t=None, sl=None, su=None):
        """Multipliers / slacks of the iterate (the other five fields of acados' load_iterate)."""
        B, N = self.B, self.N
        a = [None if v is None else _f64(v, (B, N * w)) for v, w in ((pi, 7), (lam, self.nc), (t, self.nc), (sl, 2), (su, 2))]
        check(self.L.admpc_batch_set_duals(self.h, *[_dp(v) for v in a]), "set_duals")

    def reset(self):
        check(self.L.admpc_batch_reset(self.h), "reset")

    # ---- reference generation on the device (RefTrajectory.get_waypoints for the whole batch) -----------------------
    def set_track(self, traj, H=None, traj_dt=None, anchor=False):
        """traj[L,6] = [vel, x, y, psi, cdist, curv] (the table RefTrajectory.set_traj builds); H >= N reference points.
        anchor=True measures arc length from each vehicle's closest waypoint (shared global track) instead of from the
        start of the window (the reference's literal behaviour)."""
        traj = _f64(traj)
        H = int(H if H is not None else self.N)
        dt = float(traj_dt if traj_dt is not None else self.opts.dt)
        check(self.L.admpc_batch_set_track(self.h, traj.shape[0], _dp(traj), H, dt), "set_track")
        check(self.L.admpc_batch_set_track_anchor(self.h, int(anchor)), "set_track_anchor")

    def make_yref(self):
        check(self.L.admpc_batch_make_yref(self.h), "make_yref")

    def get_yref(self):
        return self._get(self.L.admpc_batch_get_yref, self.N * 9 + 7)

    def get_waypoint_info(self):
        s0, ey, ep = (np.empty(self.B) for _ in range(3))
        stop = C.c_int()
        check(self.L.admpc_batch_get_waypoint_info(self.h, _dp(s0), _dp(ey), _dp(ep), C.byref(stop)), "get_waypoint_info")
        return s0, ey, ep, bool(stop.value)

    # ---- after-solve logic and closed loops on the device ------------------------------------------------------------
    def postsolve(self, advance=False, safe_threshold=10):
        check(self.L.admpc_batch_postsolve(self.h, int(advance), int(safe_threshold)), "postsolve")

    def closed_loop(self, steps, use_track=True, safe_threshold=10, log=False):
        """`steps` control steps on the device. Returns the plant-state log [steps+1,B,7] when log=True."""
        buf = np.empty((steps + 1, self.B, 7)) if log else None
        check(self.L.admpc_batch_closed_loop(self.h, int(steps), int(use_track), int(safe_threshold), _dp(buf)), "closed_loop")
        return buf

    def get_loop_info(self):
        valid, cnt, ok = (np.empty(self.B, dtype=np.int32) for _ in range(3))
        ua, x0 = np.empty((self.B, 2)), np.empty((self.B, 7))
        check(self.L.admpc_batch_get_loop_info(self.h, _ip(valid), _ip(cnt), _ip(ok), _dp(ua), _dp(x0)), "get_loop_info")
        return dict(valid=valid, safe_count=cnt, cmd_ok=ok, u_apply=ua, x0=x0)

    # ---- solve / outputs ------------------------------------------------------------------------------------
    def solve(self):
        check(self.L.admpc_batch_solve(self.h), "solve")

    def solve_sqp(self, max_iter=100, tol=(1e-6, 1e-6, 1e-6, 1e-6)):
        """Full SQP (nlp_solver_type "SQP", create_ros_ad_mpc.py:47-51): iterate every instance to convergence.
        Returns dict(status[B] acados codes {0,1,2,4}, sqp_iter[B], res[B,4], iterations_run)."""
        t = _f64(np.asarray(tol, dtype=np.float64).reshape(4))
        n = C.c_int(0)
        check(self.L.admpc_batch_solve_sqp(self.h, int(max_iter), _dp(t), C.byref(n)), "solve_sqp")
        st, it = np.empty(self.B, dtype=np.int32), np.empty(self.B, dtype=np.int32)
        res = np.empty((self.B, 4))
        check(self.L.admpc_batch_get_sqp_info(self.h, st.ctypes.data_as(C.POINTER(C.c_int)),
                                              it.ctypes.data_as(C.POINTER(C.c_int)), _dp(res)), "get_sqp_info")
        return dict(status=st, sqp_iter=it, res=res, iterations_run=n.value)

    def wait(self):
        check(self.L.admpc_batch_wait(self.h), "wait")

    def _get(self, fn, width):
        out = np.empty((self.B, width))
        check(fn(self.h, _dp(out)), "get")
        return out

    def get_u(self):
        return self._get(self.L.admpc_batch_get_u, self.N * 2).reshape(self.B, self.N, 2)

    def get_x(self):
        return self._get(self.L.admpc_batch_get_x, (self.N + 1) * 7).reshape(self.B, self.N + 1, 7)

    def get_pi(self):
        return self._get(self.L.admpc_batch_get_pi, self.N * 7).reshape(self.B, self.N, 7)

    def get_lam(self):
        return self._get(self.L.admpc_batch_get_lam, self.N * self.nc).reshape(self.B, self.N, self.nc)

    def get_t(self):
        return self._get(self.L.admpc_batch_get_t, self.N * self.nc).reshape(self.B, self.N, self.nc)

    def get_slacks(self):
        sl, su = np.empty((self.B, self.N * 2)), np.empty((self.B, self.N * 2))
        check(self.L.admpc_batch_get_slacks(self.h, _dp(sl), _dp(su)), "get_slacks")
        return sl.reshape(self.B, self.N, 2), su.reshape(self.B, self.N, 2)

    def get_status(self):
        st, qs, qi = (np.empty(self.B, dtype=np.int32) for _ in range(3))
        check(self.L.admpc_batch_get_status(self.h, _ip(st), _ip(qs), _ip(qi)), "get_status")
        return st, qs, qi

    def get_lin(self):
        B, N = self.B, self.N
        A, Bm, b = np.empty((B, N, 7, 7)), np.empty((B, N, 7, 2)), np.empty((B, N, 7))
        q, r = np.empty((B, N + 1, 7)), np.empty((B, N, 2))
        check(self.L.admpc_batch_get_lin(self.h, _dp(A), _dp(Bm), _dp(b), _dp(q), _dp(r)), "get_lin")
        return dict(A=A, B=Bm, b=b, q=q, r=r)

    def solve_batch(self, x0, yref, p, u_out=None, x_out=None, status_out=None):
        """One RTI step through host buffers: H2D(x0,yref,p) -> solve -> D2H(u,x,status).  Returns (u, x, status)."""
        B, N = self.B, self.N
        x0, yref = _f64(x0, (B, 7)), _f64(yref, (B, N * 9 + 7))
        p = _f64(np.broadcast_to(np.asarray(p, dtype=np.float64).reshape(-1), (B,)) if np.size(p) in (1, B) else p)
        if p.size != B:
            self.set_p(p)
            p = None
        u = u_out if u_out is not None else np.empty((B, N, 2))
        x = x_out if x_out is not None else np.empty((B, N + 1, 7))
        st = status_out if status_out is not None else np.empty(B, dtype=np.int32)
        check(self.L.admpc_batch_solve_host(self.h, _dp(x0), _dp(yref), None if p is None else _dp(p), _dp(u), _dp(x), _ip(st)), "solve_host")
        return u, x, st

    # ---- instrumentation ------------------------------------------------------------------------------------
    def set_profiling(self, on=True):
        check(self.L.admpc_batch_set_profiling(self.h, int(on)))

    def last_ms(self, name="solve"):
        ms = C.c_float()
        check(self.L.admpc_batch_last_ms(self.h, name.encode(), C.byref(ms)), "last_ms")
        return float(ms.value)

    def timer_start(self):
        check(self.L.admpc_batch_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(self.L.admpc_batch_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self):
        check(self.L.admpc_batch_flush_l2(self.h))

    def kernel_launches(self):
        return int(self.L.admpc_batch_kernel_launches(self.h))


class PipelinedSolver:
    """B instances behind ONE C handle (admpc_pipe, csrc/pipe.cu) that owns `chunks` chunk handles, one CUDA stream each:
    the H2D copy of chunk c+1, the solve of chunk c and the D2H copy of chunk c-1 overlap.  Host buffers should be pinned
    (PinnedArray) for real overlap.  `parts` are non-owning views of the chunk handles for per-chunk settings."""

    def __init__(self, B, opts=None, device=0, N=None, chunks=8):
        self.L = _lib.load()
        self.opts = opts if opts is not None else default_opts(N or 20)
        if N is not None:
            self.opts.N = N
        self.B, self.N = int(B), int(self.opts.N)
        p = C.c_void_p()
        check(self.L.admpc_pipe_create(C.byref(self.opts), self.B, int(device), int(chunks), C.byref(p)), "admpc_pipe_create")
        self.p = p
        self.parts, self.ranges = [], []
        for c in range(self.L.admpc_pipe_chunks(self.p)):
            lo, hi = C.c_int(), C.c_int()
            h = self.L.admpc_pipe_chunk(self.p, c, C.byref(lo), C.byref(hi))
            self.ranges.append((lo.value, hi.value))
            self.parts.append(BatchSolver._view(h, hi.value - lo.value, self.opts))

    def set_gp(self, model, stage0_trigger=1):
        if model is None:
            check(self.L.admpc_pipe_set_gp(self.p, 0, 0, 0, None, None, None, None, None, None, None, 0), "pipe_set_gp")
            return
        X = _f64(model["X"])
        nout, M, dz = X.shape
        feat = np.ascontiguousarray(model.get("feat", (3, 4, 5, 6)[:dz]), dtype=np.int32)
        rows = np.ascontiguousarray(model.get("rows", (4, 5)[:nout]), dtype=np.int32)
        check(self.L.admpc_pipe_set_gp(self.p, nout, M, dz, _ip(feat), _ip(rows), _dp(X), _dp(_f64(model["alpha"])),
                                       _dp(_f64(model["ell"])), _dp(_f64(model["sigma_f"])), _dp(_f64(model["y_mean"])),
                                       int(stage0_trigger)), "admpc_pipe_set_gp")

    def set_iterate(self, x=None, u=None):
        check(self.L.admpc_pipe_set_iterate(self.p, None if x is None else _dp(_f64(x, (self.B, (self.N + 1) * 7))),
                                            None if u is None else _dp(_f64(u, (self.B, self.N * 2)))), "pipe_set_iterate")

    def set_gp_ensemble(self, models, centroids=None, stage0_trigger=1):
        return [s.set_gp_ensemble(models, centroids=centroids, stage0_trigger=stage0_trigger) for s in self.parts][0]

    def select_gp(self, x=None, u=None):
        return np.concatenate([s.select_gp(None if x is None else x[lo:hi], None if u is None else u[lo:hi])
                               for s, (lo, hi) in zip(self.parts, self.ranges)])

    def set_gp_index(self, idx):
        idx = np.broadcast_to(np.asarray(idx, dtype=np.int32).reshape(-1), (self.B,))
        for s, (lo, hi) in zip(self.parts, self.ranges):
            s.set_gp_index(idx[lo:hi])

    def set_gp_state(self, gp_state):
        for s, (lo, hi) in zip(self.parts, self.ranges):
            s.set_gp_state(None if gp_state is None else gp_state[lo:hi])

    def set_kappa(self, kappa):
        """Frenet variant: kappa [B, N] (or [B])."""
        k = np.asarray(kappa, dtype=np.float64)
        for s, (lo, hi) in zip(self.parts, self.ranges):
            s.set_kappa(k[lo:hi])

    def set_kappa_spline(self, breaks=None, coef=None):
        """Frenet variant: kappa(s) spline inside the model (BatchSolver.set_kappa_spline), per instance or shared."""
        b = None if breaks is None else np.asarray(breaks, dtype=np.float64)
        c = None if coef is None else np.asarray(coef, dtype=np.float64)
        for s, (lo, hi) in zip(self.parts, self.ranges):
            if b is None:
                s.set_kappa_spline(None)
            else:
                s.set_kappa_spline(b if b.ndim == 1 else b[lo:hi], c if c.ndim == 2 else c[lo:hi])

    def wait(self):
        check(self.L.admpc_pipe_wait(self.p), "pipe_wait")

    @staticmethod
    def _c(a, dtype=np.float64):
        assert a.flags["C_CONTIGUOUS"] and a.dtype == dtype, "host buffers must be C-contiguous %s" % dtype
        return a

    def solve_batch(self, x0, yref, p, u_out, x_out, status_out):
        """x0[B,7] yref[B,N*9+7] p[B] -> u_out[B,N,2] x_out[B,N+1,7] status_out[B] (all C-contiguous float64/int32).
        One C call (admpc_pipe_solve_host): every chunk is enqueued on its stream, then all are awaited."""
        c = self._c
        check(self.L.admpc_pipe_solve_host(self.p, _dp(_f64(x0)), _dp(_f64(yref)), _dp(_f64(p)), _dp(c(u_out)), _dp(c(x_out)),
                                           _ip(c(status_out, np.int32))), "pipe_solve_host")
        return u_out, x_out, status_out

    def set_track(self, traj, H=None, traj_dt=None, anchor=False):
        traj = _f64(traj)
        H = int(H if H is not None else self.N)
        dt = float(traj_dt if traj_dt is not None else self.opts.dt)
        check(self.L.admpc_pipe_set_track(self.p, traj.shape[0], _dp(traj), H, dt, int(anchor)), "pipe_set_track")

    def solve_from_pose(self, x0, p, u_out, x_out, status_out):
        """Control step from vehicle states only: reference generation, solve and read-back per chunk, pipelined."""
        c = self._c
        check(self.L.admpc_pipe_solve_pose(self.p, _dp(_f64(x0)), _dp(_f64(p)), _dp(c(u_out)), _dp(c(x_out)),
                                           _ip(c(status_out, np.int32))), "pipe_solve_pose")
        return u_out, x_out, status_out

    def kernel_launches(self):
        return int(self.L.admpc_pipe_kernel_launches(self.p))

    def close(self):
        if getattr(self, "p", None):
            for s in self.parts:
                s.close()
            self.L.admpc_pipe_free(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AcadosOcpSolverB200:
    """Drop-in for the ``AcadosOcpSolver`` object used by AD3DOptimizer (single instance, acados-shim symbols)."""

    def __init__(self, opts=None, N=None, nlp_solver_type="SQP_RTI", nlp_solver_max_iter=100, nlp_tol=None):
        """nlp_solver_type mirrors ocp.solver_options.nlp_solver_type (ad_3d_optimizer.py:205): "SQP_RTI" or "SQP"."""
        self.L = _lib.load()
        self.c = C.c_void_p(self.L.sim_car_acados_create_capsule())
        if opts is not None:
            check(self.L.sim_car_acados_set_opts(self.c, C.byref(opts)), "set_opts")
        tol = None if nlp_tol is None else _f64(np.asarray(nlp_tol, dtype=np.float64).reshape(4))
        check(self.L.sim_car_acados_set_nlp_solver(self.c, nlp_solver_type.encode(), int(nlp_solver_max_iter),
                                                   _dp(tol) if tol is not None else None), "set_nlp_solver")
        n = N if N is not None else (opts.N if opts is not None else 40)
        check(self.L.sim_car_acados_create_with_discretization(self.c, int(n), None), "create")
        self.N = int(n)
        self.status = 0
        self.own_set = bool(opts is not None and opts.con_set == 1)      # the Frenet variant's own constraint set: 12 rows per stage

    def set(self, stage, field, value):
        v = _f64(np.atleast_1d(value)).reshape(-1)
        check(self.L.sim_car_acados_set(self.c, int(stage), field.encode(), _dp(v), v.size), "set(%s)" % field)

    def solve(self):
        self.status = self.L.sim_car_acados_solve(self.c)
        check(self.status, "solve")
        return self.status

    def get(self, stage, field):
        if self.own_set:        # ad_mpc/debug.json: 12 multipliers and 2 + 2 slacks per stage, 20 / 1 + 1 at stage 0
            n = {"x": 7, "u": 2, "pi": 7, "sl": 1 if stage == 0 else 2, "su": 1 if stage == 0 else 2}.get(field, 20 if stage == 0 else 12)
        else:
            n = {"x": 7, "u": 2, "pi": 7, "sl": 2, "su": 2}.get(field, 22 if stage == 0 else NC)
        out = np.empty(n)
        check(self.L.sim_car_acados_get(self.c, int(stage), field.encode(), _dp(out), n), "get(%s)" % field)
        return out

    def get_stats(self, name):
        if name in ("time_tot", "kkt_norm_inf"):
            v = C.c_double()
        else:
            v = C.c_int()
        check(self.L.sim_car_acados_get_stat(self.c, name.encode(), C.byref(v)), "get_stats")
        return v.value

    def print_statistics(self):
        self.L.sim_car_acados_print_stats(self.c)

    def reset(self):
        check(self.L.sim_car_acados_reset(self.c, 1), "reset")

    def store_iterate(self, filename="", overwrite=True):
        """acados store_iterate JSON layout (sim_car_iterate.json): x_k,u_k,z_k,pi_k,lam_k,t_k,sl_k,su_k."""
        d = {}
        for k in range(self.N + 1):
            d["x_%d" % k] = self.get(k, "x").tolist()
            d["z_%d" % k] = []
            for f in ("u", "lam", "t", "sl", "su"):
                d["%s_%d" % (f, k)] = self.get(k, f).tolist() if k < self.N else []
            if k < self.N:
                d["pi_%d" % k] = self.get(k, "pi").tolist()
        if filename:
            with open(filename, "w") as fh:
                json.dump(d, fh, indent=4, sort_keys=True)
        return d

    def load_iterate(self, filename):
        with open(filename) as fh:
            d = json.load(fh)
        for k in range(self.N + 1):
            self.set(k, "x", d["x_%d" % k])
            if k < self.N:
                self.set(k, "u", d["u_%d" % k])
                for f in ("pi", "lam", "t", "sl", "su"):          # acados restores all seven fields (ad_3d_optimizer.py:454)
                    if d.get("%s_%d" % (f, k)):
                        self.set(k, f, d["%s_%d" % (f, k)])

    def __del__(self):
        try:
            if self.c:
                self.L.sim_car_acados_free(self.c)
                self.L.sim_car_acados_free_capsule(self.c)
                self.c = None
        except Exception:
            pass
