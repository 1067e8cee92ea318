"""Batched host-side mirror of `AD3DOptimizer` (reference: ad_mpc/ad_3d_optimizer.py:28-480).

Same method names, argument meaning and return conventions as the reference class, with a leading batch axis on every
per-vehicle argument (B = 1 reproduces the reference object one to one).  What the reference does in Python around the
acados solver object is restated here in numpy; the solve itself is one call into libadmpc_b200 (no CPU fallback):

  __init__                 cost / bound / solver options of the OCP            (:30-209)
  set_reference_state      constant reference over the horizon                 (:312-332)
  set_reference_trajectory reference sequence, padded with its last row        (:334-383)
  run_optimization         heading unwrap of the reference relative to x0.psi  (:417-438), x0 as stage-0 bound (:441-442),
                           vel_switch blend parameter on every stage (:443-450), solve (:456), read u / x (:459-466),
                           validity check and backup control                   (:385-394, :468-476)
"""
import math

import numpy as np

from .solver import BatchSolver, default_opts


class AD3DOptimizerB200:
    def __init__(self, B=1, t_horizon=1.0, n_nodes=20, q_cost=None, r_cost=None, solver_options=None, device=0,
                 blend_min=100.0, blend_max=110.0, **opt_overrides):
        """q_cost / r_cost: diagonals of the LINEAR_LS weights (defaults of the reference, :41-45);
        solver_options: {"solver_type": "SQP_RTI" | "SQP", ...} as in the reference (:205);
        blend_min / blend_max: velocities between which vel_switch ramps from the kinematic to the dynamic model
        (ad_3d.py:62-64); opt_overrides: any other `admpc_opts` field (bounds, vehicle constants, ...)."""
        if q_cost is None:
            q_cost = np.array([10, 10, 50, 0.0, 0.0, 0.0, 1])
        if r_cost is None:
            r_cost = np.array([1.0, 100.0])
        q_cost, r_cost = np.asarray(q_cost, dtype=np.float64), np.asarray(r_cost, dtype=np.float64)
        self.B, self.T, self.N = int(B), float(t_horizon), int(n_nodes)
        self.blend_min, self.blend_max = float(blend_min), float(blend_max)
        kw = dict(dt=self.T / self.N, W=list(q_cost) + list(r_cost), We=list(q_cost * 1e-6))     # W_e = diag(q) 1e-6 (:151)
        kw.update(opt_overrides)
        self.opts = default_opts(self.N, **kw)
        self.solver_type = "SQP_RTI" if solver_options is None else solver_options.get("solver_type", "SQP_RTI")
        self.solver = BatchSolver(self.B, self.opts, device=device)
        self.target = None
        self.u_target = None
        self.x_init = None
        self.valid_present = np.zeros(self.B, dtype=bool)
        self.prev_w_opt_acados = [None] * self.B

    # ---- references ------------------------------------------------------------------------------------------
    def set_reference_state(self, x_target=None, u_target=None):
        """Constant reference [x(7)] (and [u(2)]) on every node (:312-332). x_target: [7] or [B,7]."""
        if x_target is None:
            x_target = np.zeros(7)
        if u_target is None:
            u_target = np.zeros(2)
        x = np.broadcast_to(np.asarray(x_target, dtype=np.float64).reshape(-1, 7), (self.B, 7))
        u = np.broadcast_to(np.asarray(u_target, dtype=np.float64).reshape(-1, 2), (self.B, 2))
        self.target = np.repeat(x[:, None, :], self.N + 1, axis=1)
        self.u_target = np.repeat(u[:, None, :], self.N + 1, axis=1)
        return 0

    def set_reference_trajectory(self, x_target, u_target):
        """x_target [L,7] or [B,L,7], u_target [L,2] or [B,L,2]; rows are appended (last row repeated) until N+1 (:347-349)."""
        x = np.asarray(x_target, dtype=np.float64)
        u = np.asarray(u_target, dtype=np.float64)
        if x.ndim == 2:
            x = np.broadcast_to(x[None], (self.B,) + x.shape)
        if u.ndim == 2:
            u = np.broadcast_to(u[None], (self.B,) + u.shape)
        x, u = x.copy(), u.copy()
        while x.shape[1] < self.N + 1:
            x = np.concatenate([x, x[:, -1:, :]], axis=1)
            u = np.concatenate([u, u[:, -1:, :]], axis=1)
        self.target, self.u_target = x, u
        return 0

    # ---- validity / backup (per vehicle) -----------------------------------------------------------------------
    @staticmethod
    def is_valid_command(x_opt, ref):
        """Verbatim semantics of :385-394 for one vehicle (the last entry of tmp_dist stays 0)."""
        tmp_dist = np.zeros(len(ref))
        for i in range(0, len(ref) - 1):
            tmp_dist[i] = math.sqrt((ref[i, 0] - x_opt[i, 0]) ** 2 + (ref[i, 1] - x_opt[i, 1]) ** 2)
        return bool(np.mean(tmp_dist) < 3.0 and np.cov(tmp_dist) < 2 and np.max(tmp_dist) < 4)

    # ---- per-solve marshalling (pure numpy, no device) -----------------------------------------------------------
    @staticmethod
    def marshal(x_init, target, u_target, N, blend_min, blend_max):
        """x_init [B,7], target [B,N+1,7] (its node-N heading is modified IN PLACE like the reference does, :431-437),
        u_target [B,>=N,2] -> (yref [B, 9N+7], vel_switch [B]).  Vectorised form of ad_3d_optimizer.py:417-443."""
        B = x_init.shape[0]
        ref = np.concatenate([target[:, :N, :], u_target[:, :N, :]], axis=2)       # [B, N, 9]
        psi0 = x_init[:, 2]
        # heading unwrap relative to the vehicle heading (:423-428), per vehicle and node
        neg, pos = (psi0 < 0)[:, None], (psi0 > 0)[:, None]
        r2 = ref[:, :, 2].copy()
        out = np.where(neg & (psi0[:, None] + math.pi < r2), r2 - 2 * math.pi, r2)
        out = np.where(pos & (psi0[:, None] - math.pi > r2), r2 + 2 * math.pi, out)
        ref[:, :, 2] = out
        tN = target[:, N, 2].copy()
        tN_new = np.where((psi0 < 0) & (psi0 + math.pi < tN), tN - 2 * math.pi, tN)
        tN_new = np.where((psi0 > 0) & (psi0 - math.pi > tN), tN + 2 * math.pi, tN_new)
        target[:, N, 2] = tN_new
        yref = np.concatenate([ref.reshape(B, N * 9), target[:, N, :]], axis=1)
        vel_switch = np.clip((x_init[:, 3] - blend_min) / (blend_max - blend_min), 0.0, 1.0)   # :443
        return yref, vel_switch

    # ---- the solve ----------------------------------------------------------------------------------------------
    def run_optimization(self, initial_state=None, use_model=0, return_x=False, gp_regression_state=None):
        """initial_state [7] or [B,7]; returns the flattened control sequence(s) [B, 2N] (B = 1: [2N]) and, with
        return_x, (w_opt, x_opt [B,N+1,7], solver_status [B]) like the reference (:396-480)."""
        if initial_state is None:
            initial_state = np.zeros(7)
        x_init = np.asarray(initial_state, dtype=np.float64).reshape(-1, 7)
        x_init = np.broadcast_to(x_init, (self.B, 7)).copy()
        self.x_init = x_init
        if self.target is None:
            self.set_reference_state()
        N = self.N
        tgt = self.target
        yref, vel_switch = self.marshal(x_init, tgt, self.u_target, N, self.blend_min, self.blend_max)
        s = self.solver
        s.set_x0(x_init)
        s.set_yref(yref)
        s.set_p(vel_switch)
        if gp_regression_state is not None:
            s.set_gp_state(np.broadcast_to(np.asarray(gp_regression_state, dtype=np.float64).reshape(-1, 7), (self.B, 7)))
        if self.solver_type == "SQP":
            info = s.solve_sqp()
            solver_status = info["status"]
        else:
            s.solve()
            solver_status = s.get_status()[0]
        u = s.get_u()
        x = s.get_x()
        w_opt = u.reshape(self.B, N * 2).copy()
        for b in range(self.B):
            if self.is_valid_command(x[b], tgt[b]):
                self.valid_present[b] = True
                self.prev_w_opt_acados[b] = w_opt[b].copy()
            elif self.prev_w_opt_acados[b] is not None:
                prev = self.prev_w_opt_acados[b]
                w_opt[b] = self._backup(prev)
        if self.B == 1:
            return w_opt[0] if not return_x else (w_opt[0], x[0], int(solver_status[0]))
        return w_opt if not return_x else (w_opt, x, solver_status)

    @staticmethod
    def _backup(prev):
        """`np.concatenate((prev[2:-1], prev[-3:-1]))` (:472): the previous sequence shifted by one node; the reference's
        slice arithmetic yields 2N-1 entries (the caller only reads the first pair), padded here to 2N with the last."""
        w = np.concatenate((prev[2:-1], prev[-3:-1]))
        return np.concatenate((w, w[-1:]))

    def set_gp(self, model, stage0_trigger=1):
        self.solver.set_gp(model, stage0_trigger=stage0_trigger)

    def close(self):
        self.solver.close()
