"""Synthetic workloads of SURVEY.md 8(d): circular reference track, perturbed initial states, random-init GP.

Host-side numpy only.  The reference has no dataset / fitted GP in-tree, so benchmarks and
parity tests run on these generators.  psi-unwrapping of the reference follows
data_driven_mpc/ros_gp_mpc/src/ad_mpc/ad_3d_optimizer.py:423-437.
"""
import math

import numpy as np

TRACK_R = 50.0
V_REF = 8.0
X0_SIGMA = np.array([0.3, 0.3, 0.05, 0.5, 0.1, 0.05, 0.02])
GP_BOX_LO = np.array([2.0, -1.0, -0.8, -0.52])
GP_BOX_HI = np.array([14.0, 1.0, 0.8, 0.52])


def unwrap_ref_psi(psi0, psi_ref):
    """ad_3d_optimizer.py:423-437: shift the reference heading by 2*pi towards the current heading."""
    psi_ref = np.array(psi_ref, dtype=np.float64, copy=True)
    if np.ndim(psi0) == 0:
        psi0 = np.full(psi_ref.shape[:1], float(psi0))
    psi0 = np.asarray(psi0, dtype=np.float64).reshape(-1, *([1] * (psi_ref.ndim - 1)))
    neg = (psi0 < 0) & (psi0 + math.pi < psi_ref)
    pos = (psi0 > 0) & (psi0 - math.pi > psi_ref)
    psi_ref[neg] -= 2 * math.pi
    psi_ref[pos] += 2 * math.pi
    return psi_ref


def circle_reference(theta0, N, dt, v_ref=V_REF, R=TRACK_R):
    """ref_k = (R cos th_k, R sin th_k, th_k + pi/2, v_ref, 0, 0, 0), th_k = th_0 + k v dt / R.  -> [B, N+1, 7]"""
    theta0 = np.atleast_1d(np.asarray(theta0, dtype=np.float64))
    k = np.arange(N + 1)
    th = theta0[:, None] + k[None, :] * v_ref * dt / R
    ref = np.zeros((theta0.shape[0], N + 1, 7))
    ref[..., 0] = R * np.cos(th)
    ref[..., 1] = R * np.sin(th)
    psi = th + math.pi / 2
    ref[..., 2] = (psi + math.pi) % (2 * math.pi) - math.pi       # headings arrive wrapped to [-pi, pi)
    ref[..., 3] = v_ref
    return ref


def make_batch(B, N, dt=0.05, seed=20262, p=0.0, perturb=1.0):
    """Batch of independent MPC instances on the circular track (SURVEY 8d cfg 2/3).

    Returns dict: x0[B,7], yref[B,N*9+7] (stage rows [x_ref(7),u_ref(2)] + terminal x_ref), p[B,N],
    x_init[B,N+1,7], u_init[B,N,2] (warm iterate = reference rollout, u=0), ref[B,N+1,7].
    """
    rng = np.random.default_rng(seed)
    theta0 = rng.uniform(0.0, 2 * math.pi, size=B)
    ref = circle_reference(theta0, N, dt)
    x0 = ref[:, 0, :] + perturb * rng.normal(size=(B, 7)) * X0_SIGMA
    x0[:, 2] = (x0[:, 2] + math.pi) % (2 * math.pi) - math.pi
    ref[..., 2] = unwrap_ref_psi(x0[:, 2], ref[..., 2])
    yref = np.zeros((B, N * 9 + 7))
    yref[:, :N * 9].reshape(B, N, 9)[:, :, :7] = ref[:, :N, :]
    yref[:, N * 9:] = ref[:, N, :]
    x_init = ref.copy()
    u_init = np.zeros((B, N, 2))
    return dict(x0=x0, yref=yref, p=np.full((B, N), float(p)), x_init=x_init, u_init=u_init, ref=ref)


def _sqexp(Xa, Xb, ell, sigma_f):
    d = (Xa[:, None, :] - Xb[None, :, :]) / ell
    return sigma_f * np.exp(-0.5 * np.sum(d * d, axis=2))


def make_gp(M=200, seed=20263, n_out=2, dz=4, sigma_f=0.5, sigma_n=0.01):
    """'Random-init' GP of SURVEY 8(d): per output, X ~ U(box), ell ~ U[0.5,2]*box half-width, alpha = K^-1 (y - ybar)
    by Cholesky in FP64 (fit formulas: model_fitting/gp.py:361-363)."""
    rng = np.random.default_rng(seed)
    lo, hi = GP_BOX_LO[:dz], GP_BOX_HI[:dz]
    X = np.zeros((n_out, M, dz))
    alpha = np.zeros((n_out, M))
    ell = np.zeros((n_out, dz))
    y_mean = np.zeros(n_out)
    for j in range(n_out):
        X[j] = rng.uniform(lo, hi, size=(M, dz))
        ell[j] = rng.uniform(0.5, 2.0, size=dz) * 0.5 * (hi - lo)
        vx, vy, dl = X[j][:, 0], X[j][:, min(1, dz - 1)], X[j][:, min(3, dz - 1)]
        y = (0.3 * np.sin(vy) + 0.1 * dl * vx / 10.0) * (1.0 if j == 0 else 0.5) + rng.normal(size=M) * sigma_n
        y_mean[j] = y.mean()
        K = _sqexp(X[j], X[j], ell[j], sigma_f) + sigma_n ** 2 * np.eye(M)
        L = np.linalg.cholesky(K)
        alpha[j] = np.linalg.solve(L.T, np.linalg.solve(L, y - y_mean[j]))
    return dict(X=X, alpha=alpha, ell=ell, sigma_f=np.full(n_out, sigma_f), y_mean=y_mean,
                feat=(3, 4, 5, 6)[:dz], rows=(4, 5)[:n_out])


def make_batch_frenet(B, N, dt=0.05, seed=20270, p=1.0, perturb=1.0, radius=50.0, v_ref=8.0):
    """Frenet-variant twin of make_batch (SURVEY 8a A2'): states [s, e_y, e_psi, v_x, v_y, r, delta] along a circular
    path of the given radius (curvature 1/R at every node), reference = on the path at v_ref.
    Returns the make_batch dict plus kappa[B, N]."""
    rng = np.random.default_rng(seed)
    s0 = rng.uniform(0.0, 2 * math.pi * radius, size=B)
    ref = np.zeros((B, N + 1, 7))
    ref[..., 0] = s0[:, None] + v_ref * dt * np.arange(N + 1)[None, :]
    ref[..., 3] = v_ref
    x0 = ref[:, 0, :] + perturb * rng.normal(size=(B, 7)) * X0_SIGMA
    yref = np.zeros((B, N * 9 + 7))
    yref[:, :N * 9].reshape(B, N, 9)[:, :, :7] = ref[:, :N, :]
    yref[:, N * 9:] = ref[:, N, :]
    pp = np.full((B, N), float(p))
    kappa = np.full((B, N), 1.0 / radius) * (1.0 + 0.2 * rng.uniform(-1, 1, size=(B, 1)))
    return dict(x0=x0, yref=yref, p=pp, x_init=ref.copy(), u_init=np.zeros((B, N, 2)), ref=ref, kappa=kappa)
