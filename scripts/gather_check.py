"""2-rank check of admpc_batch_gather (run under torchrun on 2 GPUs): gathered blocks == per-rank results."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from ad_mpc_b200 import BatchSolver, default_opts, workload as wl, _lib
dist.init_process_group("gloo")
rank, world, local = dist.get_rank(), dist.get_world_size(), int(os.environ["LOCAL_RANK"])
L = _lib.load()
B, N = 2048, 20
s = BatchSolver(B, default_opts(N), device=local)
uid = (C.c_char * 128)()
if rank == 0: _lib.check(L.admpc_nccl_unique_id(uid))
obj = [bytes(uid.raw)]; dist.broadcast_object_list(obj, src=0)
uid = (C.c_char * 128).from_buffer_copy(obj[0])
_lib.check(L.admpc_batch_comm_init(s.h, uid, rank, world), "comm_init")
batch = wl.make_batch(B, N, seed=100 + rank, p=1.0)
s.set_iterate(batch["x_init"], batch["u_init"]); s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"][:, 0])
s.solve()
u_all = np.zeros((world * B, N, 2)); x_all = np.zeros((world * B, N + 1, 7)); st_all = np.full(world * B, -1, dtype=np.int32)
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
for rep in range(3):
    s.timer_start()
    _lib.check(L.admpc_batch_gather(s.h, 0, None, None, None), "gather")
    ms = s.timer_stop()
_lib.check(L.admpc_batch_gather(s.h, 0, u_all.ctypes.data_as(dp), x_all.ctypes.data_as(dp), st_all.ctypes.data_as(ip)), "gather")
mine = (s.get_u(), s.get_x(), s.get_status()[0])
blocks = [None] * world if rank == 0 else None
dist.gather_object(mine, blocks, dst=0)
if rank == 0:
    for r in range(world):
        assert np.array_equal(u_all[r * B:(r + 1) * B], blocks[r][0]) and np.array_equal(x_all[r * B:(r + 1) * B], blocks[r][1])
        assert np.array_equal(st_all[r * B:(r + 1) * B], blocks[r][2])
    print("GATHER_OK world=%d device-gather %.3f ms" % (world, ms))
# the blocks left on the root's device can be fetched again without another transfer
if rank == 0:
    u2 = np.zeros_like(u_all); x2 = np.zeros_like(x_all); st2 = np.full_like(st_all, -1)
    _lib.check(L.admpc_batch_get_gathered(s.h, u2.ctypes.data_as(dp), x2.ctypes.data_as(dp), st2.ctypes.data_as(ip)), "get_gathered")
    assert np.array_equal(u2, u_all) and np.array_equal(x2, x_all) and np.array_equal(st2, st_all)
    print("GET_GATHERED_OK")
# ---- fused gather: the QP kernel's epilogue writes straight into the root's block (CUDA IPC peer memory) ------------------
mode = L.admpc_batch_gather_enable(s.h, 0)
_lib.check(mode, "gather_enable")
def check(tag):
    u_all[:] = 0; x_all[:] = 0; st_all[:] = -1
    _lib.check(L.admpc_batch_gather(s.h, 0, u_all.ctypes.data_as(dp), x_all.ctypes.data_as(dp), st_all.ctypes.data_as(ip)), "gather")
    mine = (s.get_u(), s.get_x(), s.get_status()[0])
    blocks = [None] * world if rank == 0 else None
    dist.gather_object(mine, blocks, dst=0)
    if rank == 0:
        for r in range(world):
            assert np.array_equal(u_all[r * B:(r + 1) * B], blocks[r][0]), tag
            assert np.array_equal(x_all[r * B:(r + 1) * B], blocks[r][1]), tag
            assert np.array_equal(st_all[r * B:(r + 1) * B], blocks[r][2]), tag
        print("FUSED_GATHER_OK mode=%d %s" % (mode, tag))
batch = wl.make_batch(B, N, seed=300 + rank, p=1.0)
s.set_iterate(batch["x_init"], batch["u_init"]); s.set_x0(batch["x0"]); s.set_yref(batch["yref"])
s.solve(); check("after solve (written by the kernel epilogue)")
s.solve(); s.solve(); check("after two more RTI steps")
s.set_iterate(batch["x_init"], batch["u_init"]); check("after set_iterate (explicit pack into the peer block)")
x0bad = batch["x0"].copy(); x0bad[3, 0] = np.nan          # one NaN instance: early-exit path of the kernel
s.set_x0(x0bad); s.solve(); check("with a failed instance")
for rep in range(3):
    s.set_x0(batch["x0"]); s.set_iterate(batch["x_init"], batch["u_init"])
    s.timer_start(); s.solve(); _lib.check(L.admpc_batch_gather(s.h, 0, None, None, None), "gather"); ms = s.timer_stop()
if rank == 0: print("solve+gather %.3f ms (mode %d)" % (ms, mode))
s.close(); dist.destroy_process_group()
