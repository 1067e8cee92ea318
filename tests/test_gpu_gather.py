"""Fused solution gather on ONE GPU: a one-rank communicator makes this process the root, so the QP
kernels' epilogue writes its [u | x | status] block into the local gathered buffer -- the same stores that go over NVLink
peer memory on a multi-GPU box (scripts/gather_check.py covers two ranks).  The gathered block must equal the iterate."""
import ctypes as C

import numpy as np
import pytest

from ad_mpc_b200 import BatchSolver, default_opts, workload as wl, _lib

pytestmark = pytest.mark.gpu


def _one_rank_comm(s):
    L = _lib.load()
    uid = (C.c_char * 128)()
    if L.admpc_nccl_unique_id(uid) != 0:
        pytest.skip("NCCL library not loadable in this process")
    _lib.check(L.admpc_batch_comm_init(s.h, uid, 0, 1), "comm_init")
    mode = L.admpc_batch_gather_enable(s.h, 0)
    _lib.check(mode, "gather_enable")
    return L, mode


def _gather(L, s, B, N):
    u = np.zeros((B, N, 2)); x = np.zeros((B, N + 1, 7)); st = np.full(B, -7, dtype=np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    _lib.check(L.admpc_batch_gather(s.h, 0, u.ctypes.data_as(dp), x.ctypes.data_as(dp), st.ctypes.data_as(ip)), "gather")
    return u, x, st


@pytest.mark.parametrize("B,N,variant", [(37, 20, 0), (33, 40, 0), (5, 90, 0), (37, 20, 1), (9, 40, 1)])
def test_fused_gather_equals_iterate(B, N, variant):
    frenet = variant == 1
    batch = (wl.make_batch_frenet if frenet else wl.make_batch)(B, N, seed=900 + B + N, p=1.0)
    s = BatchSolver(B, default_opts(N, model_variant=variant))
    L, mode = _one_rank_comm(s)
    assert mode == 1
    u_init = batch["u_init"].copy()
    x0 = batch["x0"].copy()
    if B > 4:
        u_init[3, 2, 0] = np.nan               # NaN linearisation: this instance leaves the QP kernel through the early exit
        x0[4, 0] = np.nan                      # NaN in the QP data: solved and failed (status 4), iterate untouched
    s.set_iterate(batch["x_init"], u_init)
    s.set_x0(x0); s.set_yref(batch["yref"]); s.set_p(batch["p"])
    if frenet:
        s.set_kappa(batch["kappa"])
    for step in range(2):
        s.solve()
        u, x, st = _gather(L, s, B, N)
        assert np.array_equal(u, s.get_u(), equal_nan=True) and np.array_equal(x, s.get_x(), equal_nan=True)
        assert np.array_equal(st, s.get_status()[0])
    if B > 4:
        assert st[3] == 1 and st[4] == 4 and (np.delete(st, [3, 4]) == 0).all()
    # iterate replaced after the solve: the call packs the block explicitly
    s.set_iterate(batch["x_init"], batch["u_init"])
    u, x, st = _gather(L, s, B, N)
    assert np.array_equal(u, batch["u_init"]) and np.array_equal(x, batch["x_init"])
    s.close()
