"""Host <-> device copy probe for the end-to-end path (measurement helper, not product; uses torch for pinned buffers and streams).
Every rank moves what one `admpc_pipe_solve_host` call moves at the benchmark configuration (25.6 MB up, 24.6 MB down, pinned
host memory, 8 chunks on alternating streams, both directions in flight at once) with NO compute in between, all ranks at the same
time.  If this alone takes as long as the end-to-end call loses against the device-resident step, the loss is the host side of the
box (PCIe / host memory), not the solver.
usage: torchrun --nproc-per-node N scripts/copy_probe.py   (or plain python for one GPU)"""
import os, statistics, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
H2D, D2H, CH = 25559040, 24576000, 8
hin = torch.empty(H2D, dtype=torch.uint8).pin_memory(); hout = torch.empty(D2H, dtype=torch.uint8).pin_memory()
din = torch.empty(H2D, dtype=torch.uint8, device="cuda"); dout = torch.zeros(D2H, dtype=torch.uint8, device="cuda")
streams = [torch.cuda.Stream() for _ in range(CH)]
def call(mode):
    for c, st in enumerate(streams):
        a0, a1 = c * H2D // CH, (c + 1) * H2D // CH
        b0, b1 = c * D2H // CH, (c + 1) * D2H // CH
        with torch.cuda.stream(st):
            if mode in ("both", "h2d"):
                din[a0:a1].copy_(hin[a0:a1], non_blocking=True)
            if mode in ("both", "d2h"):
                hout[b0:b1].copy_(dout[b0:b1], non_blocking=True)
    torch.cuda.synchronize()
res = {}
for mode in ("h2d", "d2h", "both"):
    ts = []
    for it in range(25):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        call(mode)
        if it >= 5:
            ts.append((time.perf_counter() - t0) * 1e3)
    res[mode] = (statistics.median(ts), max(ts))
out = [None] * world
if world > 1:
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    for mode, nbytes in (("h2d", H2D), ("d2h", D2H), ("both", H2D + D2H)):
        p50 = max(o[mode][0] for o in out)
        print("copy_probe ranks=%d %-4s p50 (slowest rank) %.3f ms  = %.1f GB/s per GPU, %.1f GB/s aggregate" % (
            world, mode, p50, nbytes / p50 / 1e6, world * nbytes / p50 / 1e6), flush=True)
if world > 1:
    dist.destroy_process_group()
