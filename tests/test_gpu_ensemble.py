"""GP ensembles (GPEnsemble, model_fitting/gp.py:536-770): K cluster models, nearest-centroid choice per instance,
one model per instance and solve.  Oracle = the single-model CPU oracle run per group of instances that chose the same
cluster, the choice itself = the reference's one-line argmin (gp.py:770) in numpy."""
import numpy as np
import pytest

from ad_mpc_b200 import BatchSolver, default_opts, workload as wl
from oracle import oracle as orc
from util_parity import mirror_opts, mixed_err

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _select_oracle(z, centroids):
    """gp.py:770: np.argmin(np.sqrt(np.sum((z[np.newaxis, :, :] - centroids[:, :, np.newaxis]) ** 2, 1)), 0), z [dz, n]."""
    return np.argmin(np.sqrt(np.sum((z[np.newaxis, :, :] - centroids[:, :, np.newaxis]) ** 2, 1)), 0)


@pytest.mark.parametrize("K,M,variant", [(3, 60, 0), (5, 33, 0), (2, 40, 1)])
def test_ensemble_selection_and_step_parity(K, M, variant):
    B, N = 90, 20
    batch = (wl.make_batch_frenet if variant else wl.make_batch)(B, N, seed=600 + K, p=1.0, perturb=3.0)
    models = [wl.make_gp(M=M, seed=40 + c) for c in range(K)]
    rng = np.random.default_rng(K)
    centroids = np.stack([rng.uniform(wl.GP_BOX_LO[:4], wl.GP_BOX_HI[:4]) for _ in range(K)])
    centroids[:, 0] = np.linspace(4.0, 12.0, K)[::-1]            # unsorted on purpose: set_gp_ensemble sorts
    opts = default_opts(N, model_variant=variant)
    s = BatchSolver(B, opts)
    order = s.set_gp_ensemble(models, centroids=centroids)
    models_s, cen_s = [models[i] for i in order], centroids[order]
    assert (np.diff(cen_s[:, 0]) >= 0).all()
    s.set_iterate(batch["x_init"], batch["u_init"])
    s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"])
    if variant:
        s.set_kappa(batch["kappa"])
    s.set_gp_state(batch["x0"])
    # choice from the current x0 (features v_x, v_y, r, delta = state indices 3..6)
    sel = s.select_gp()
    ref_sel = _select_oracle(batch["x0"][:, 3:7].T, cen_s)
    assert np.array_equal(sel, ref_sel) and len(set(sel.tolist())) >= 2
    # choice from an explicit query state (the reference queries with the reference state)
    xq = batch["ref"][:, N // 2, :]
    sel2 = s.select_gp(x=xq)
    assert np.array_equal(sel2, _select_oracle(xq[:, 3:7].T, cen_s))
    s.set_gp_index(sel)
    s.solve()
    u, x = s.get_u(), s.get_x()
    st, qs, qi = s.get_status()
    o = mirror_opts(opts)
    for c in range(K):
        idx = np.where(sel == c)[0]
        if idx.size == 0:
            continue
        gp = orc.Gp(models_s[c])
        oc = mirror_opts(opts)
        gp.apply(oc, feat=models_s[c]["feat"], rows=models_s[c]["rows"])
        r = orc.rti_batch(oc, batch["x0"][idx], batch["yref"][idx], batch["p"][idx], batch["x_init"][idx], batch["u_init"][idx],
                          gp=gp, gp_state=batch["x0"][idx], kappa=batch["kappa"][idx] if variant else None)
        assert np.array_equal(st[idx], r["status"]) and np.array_equal(qi[idx], r["qp_iter"])
        assert mixed_err(u[idx], r["u"]) <= TOL and mixed_err(x[idx], r["x"]) <= TOL
    s.close()


def test_single_model_is_ensemble_of_one_and_index_validation():
    from ad_mpc_b200 import _lib
    B, N = 16, 20
    batch = wl.make_batch(B, N, seed=610, p=1.0)
    model = wl.make_gp(M=30, seed=1)
    a = BatchSolver(B, default_opts(N)); a.set_gp(model)
    b = BatchSolver(B, default_opts(N)); b.set_gp_ensemble([model], centroids=np.zeros((1, 4)))
    outs = []
    for s in (a, b):
        s.set_iterate(batch["x_init"], batch["u_init"]); s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"])
        assert (s.select_gp() == 0).all()
        s.solve()
        outs.append(s.get_u())
    assert np.array_equal(outs[0], outs[1])
    with pytest.raises(_lib.AdmpcError):
        b.set_gp_index(np.full(B, 1))
    a.close(); b.close()
