#!/usr/bin/env python
"""bench.py -- SQP-RTI MPC solves/sec (N=20, GP-augmented) on 1/2/4/8 B200, roofline + CPU baseline.

  python bench.py [--gpus N --steps K --warmup W]                 our arm (CUDA, through the C ABI)
  python bench.py --impl reference [...]                          reference arm: CPU restatement on all host cores
  torchrun --nproc-per-node N bench.py --gpus N ...               one rank per GPU (weak scaling: B per GPU fixed)

One "step" = one SQP-RTI iteration (prepare -> QP -> update) over one batch of B instances per GPU.
Workload (BASELINE.json configs[2], SURVEY 8d cfg 3): GP-augmented bicycle NMPC, N=20, dt=0.05, RBF GP with M=200
training points on (v_y, yaw-rate), B=16384 instances per GPU, dynamic tyre model (p=1), circular track, synthetic.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(N=20, dt=0.05, M=200, B=16384, p=1.0, dz=4, n_out=2)


def algorithmic_ops(N, M, dz, n_out, n_ipm, gp=True):
    """SURVEY.md 8(d) op model (sin/cos/exp/div = 1 op): per solve, split by kernel."""
    f_sim = 4848.0 * N + 60.0 * N
    f_gp = 4.0 * N * (n_out * M * (6 * dz + 4) + 144) if gp else 0.0
    f_qp = 3200.0 * N * n_ipm
    return dict(prepare=f_sim + f_gp, qp=f_qp)


def algorithmic_bytes(N):
    """x0(7)+yref(9N+7)+p(1) in, u(2N)+x(7N+7) out, 8 B status/iter."""
    return 8 * (7 + 9 * N + 7 + 1 + 2 * N + 7 * N + 7) + 8


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        # median over the busiest half of the samples (the region also contains host-side gaps)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return dict(sm_mhz=statistics.median(load) if load else None, sm_max_mhz=smax, reasons=sorted(reasons),
                    samples=len(sm))


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def _workload_module():
    """ad_mpc_b200/workload.py is numpy-only; load it by path so that the reference arm never imports the product package
    (and therefore never loads libadmpc_b200.so)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("admpc_workload", os.path.join(ROOT, "ad_mpc_b200", "workload.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_workload(rank, B=None):
    wl = _workload_module()
    B = B or CFG["B"]
    batch = wl.make_batch(B, CFG["N"], dt=CFG["dt"], seed=20263 + 1000 * rank, p=CFG["p"])
    model = wl.make_gp(M=CFG["M"], seed=20263, n_out=CFG["n_out"], dz=CFG["dz"])
    return batch, model


def config_dict(world):
    return {"workload": "cfg3: GP-augmented bicycle NMPC (in-tree Cartesian-pose model), SQP-RTI, N=20, dt=0.05, RBF GP M=200 (d_z=4, "
                        "2 outputs on v_y/yaw-rate), dynamic tyre model p=1, B=16384 instances per GPU, circular track R=50 m",
            "N": CFG["N"], "gp_points": CFG["M"], "batch_per_gpu": CFG["B"], "global_batch": CFG["B"] * world,
            "parallelism": "dp%d (independent instances, weak scaling)" % world,
            "l2": "flushed between timed steps (256 MiB fill, untimed); per-step working set ~0.6 GB > 126 MB L2"}


# ------------------------------------------------------------------------------------------ CPU baseline -------
_ORC_FLAGS = None


def _oracle():
    """The CPU restatement, rebuilt -O3 -march=native on this machine; test infrastructure, timed here as
    the reported CPU baseline only."""
    global _ORC_FLAGS
    from oracle import oracle as orc
    if _ORC_FLAGS is None:
        _ORC_FLAGS = orc.use_native()
    return orc


def _oracle_opts(orc, N):
    """Oracle options = its own defaults (tests/test_host_cpu.py checks they equal the product's field by field)."""
    o = orc.default_opts(N=N)
    o.dt = CFG["dt"]
    return o


def cpu_oracle_rate(sample, nthreads=0, gp=True, reps=1, N=None, M=None, p=None):
    """Times the oracle port (CPU restatement of the reference path; acados itself is not buildable here) on a
    bounded sample of the workload.  Returns (solves/s, seconds, threads)."""
    orc = _oracle()
    wl = _workload_module()
    N = N or CFG["N"]
    batch = wl.make_batch(sample, N, dt=CFG["dt"], seed=20263, p=CFG["p"] if p is None else p)
    o = _oracle_opts(orc, N)
    g = None
    if gp:
        model = wl.make_gp(M=M or CFG["M"], seed=20263, n_out=CFG["n_out"], dz=CFG["dz"])
        g = orc.Gp(model)
        g.apply(o, feat=model["feat"], rows=model["rows"])
    cores = nthreads or (os.cpu_count() or 1)
    w = min(64, sample)
    orc.rti_batch(o, batch["x0"][:w], batch["yref"][:w], batch["p"][:w], batch["x_init"][:w], batch["u_init"][:w],
                  gp=g, nthreads=cores)   # warm-up (page-in, thread pool)
    t0 = time.perf_counter()
    for _ in range(reps):
        r = orc.rti_batch(o, batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], gp=g, nthreads=cores)
    dt = time.perf_counter() - t0
    assert (r["status"] == 0).all()
    return sample * reps / dt, dt, cores


def cpu_single_latency(steps=50):
    """cfg 1 on ONE host thread: nominal model, N=20, B=1, cold iterate then `steps` closed-loop RTI steps (plant =
    model prediction), per-step latency of the CPU restatement (the quantity the GPU single-instance number sits beside)."""
    orc = _oracle()
    wl = _workload_module()
    N = CFG["N"]
    b1 = wl.make_batch(1, N, seed=20261, p=0.0)
    o = _oracle_opts(orc, N)
    x0, xi, ui = b1["x0"].copy(), b1["x_init"].copy(), b1["u_init"].copy()
    ms = []
    for it in range(5 + steps):
        t0 = time.perf_counter()
        r = orc.rti_batch(o, x0, b1["yref"], b1["p"], xi, ui, nthreads=1)
        ms.append((time.perf_counter() - t0) * 1e3)
        xi, ui = r["x"], r["u"]
        x0 = r["x"][:, 1, :].copy()
    ms = sorted(ms[5:])
    return ms[len(ms) // 2], ms[min(len(ms) - 1, int(0.99 * len(ms)))]


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    sample = CFG["B"]                        # the full per-GPU batch of the config, every step
    cores = os.cpu_count() or 1
    times = []
    for s in range(args.warmup + args.steps):
        rate, dt, cores = cpu_oracle_rate(sample)
        if s >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms * 1e-3)
    line = {"impl": "reference", "metric": "SQP-RTI MPC solves/sec (N=20, GP-augmented)", "value": value, "unit": "solves/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(world),
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": "the full %d-instance cfg3 batch per step, OpenMP over instances, oracle built %s "
                                       "(CPU restatement of the acados path; acados/HPIPM are un-vendored and cannot be "
                                       "built here)" % (sample, _ORC_FLAGS)},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bind_numa(local):
    """Best effort: pin this rank's host thread (and therefore its pinned allocations, first touch) to the NUMA node of
    its GPU.  Returns a short description for the JSON line."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "gpu numa node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return "node %d has no allowed cpu (allowed %d cpus)" % (node, len(allowed))
        os.sched_setaffinity(0, use)
        return "node %d, %d cpus" % (node, len(use))
    except Exception as e:           # noqa: BLE001 -- reporting only
        return "not bound (%s)" % type(e).__name__


# ------------------------------------------------------------------------------------------ our arm ------------
def run_ours(args):
    rank, world, local = dist_env()
    from ad_mpc_b200 import BatchSolver, PinnedArray, default_opts, _lib
    import ctypes as C

    # keep NCCL's banner / debug lines off stdout: rank 0 prints exactly one JSON line there
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist = None
    if world > 1:
        import torch.distributed as dist          # plumbing only: rendezvous, barrier, max over ranks
        dist.init_process_group(backend="gloo")
    L = _lib.load()
    numa = bind_numa(local)
    B, N = CFG["B"], CFG["N"]
    opts = default_opts(N)
    opts.dt = CFG["dt"]
    s = BatchSolver(B, opts, device=local)
    batch, model = make_workload(rank)

    if world > 1:
        # NCCL communicator of the library itself: unique id from rank 0, shipped through the launcher's store
        uid = (C.c_char * 128)()
        if rank == 0:
            _lib.check(L.admpc_nccl_unique_id(uid), "nccl_unique_id")
        obj = [bytes(uid.raw)]
        dist.broadcast_object_list(obj, src=0)
        uid = (C.c_char * 128).from_buffer_copy(obj[0])
        _lib.check(L.admpc_batch_comm_init(s.h, uid, rank, world), "comm_init")
        gather_mode = L.admpc_batch_gather_enable(s.h, 0)       # 1: fused peer-memory gather, 0: NCCL send/recv
        _lib.check(gather_mode, "gather_enable")
        # GP model lives on rank 0 and is broadcast over NVLink
        X = np.ascontiguousarray(model["X"]); al = np.ascontiguousarray(model["alpha"]); ell = np.ascontiguousarray(model["ell"])
        sf = np.ascontiguousarray(model["sigma_f"]); ym = np.ascontiguousarray(model["y_mean"])
        feat = np.ascontiguousarray(model["feat"], dtype=np.int32); rows = np.ascontiguousarray(model["rows"], dtype=np.int32)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _lib.check(L.admpc_batch_bcast_gp(s.h, 0, X.shape[0], X.shape[1], X.shape[2], feat.ctypes.data_as(ip),
                                          rows.ctypes.data_as(ip), X.ctypes.data_as(dp), al.ctypes.data_as(dp),
                                          ell.ctypes.data_as(dp), sf.ctypes.data_as(dp), ym.ctypes.data_as(dp), 1), "bcast_gp")
    else:
        s.set_gp(model)

    def barrier():
        s.wait()
        if dist is not None:
            dist.barrier()

    def gather_device():
        if world > 1:
            _lib.check(L.admpc_batch_gather(s.h, 0, None, None, None), "gather")

    # resident inputs
    s.set_x0(batch["x0"]); s.set_yref(batch["yref"]); s.set_p(batch["p"][:, 0])
    x_init, u_init = batch["x_init"], batch["u_init"]

    def device_step():
        s.solve()
        gather_device()

    # ---- device-resident throughput ("value") -----------------------------------------------------------------
    s.set_profiling(True)
    sampler = ClockSampler(local)
    sampler.start()                          # sampled from the warm-up through the timed steps (GPU continuously busy)
    for _ in range(max(args.warmup, 3) + 20):
        s.set_iterate(x_init, u_init)
        device_step()
    s.wait()
    barrier()
    wall0 = time.perf_counter()
    step_ms, prep_ms, qp_ms = [], [], []
    launches0 = s.kernel_launches()
    counted = 0
    for _ in range(args.steps):
        s.set_iterate(x_init, u_init)       # same warm start every step (untimed restore of the iterate)
        s.flush_l2()
        if world > 1:
            # untimed alignment: the fused gather ends in an all-reduce, so without it a step would also time the other
            # ranks' (untimed) restore / L2-flush work of the previous iteration
            _lib.check(L.admpc_batch_barrier(s.h), "barrier")
        l0 = s.kernel_launches()
        s.timer_start()
        device_step()
        step_ms.append(s.timer_stop())
        counted += s.kernel_launches() - l0
        prep_ms.append(s.last_ms("prepare"))
        qp_ms.append(s.last_ms("qp"))
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    st, qs, qi = s.get_status()
    assert (st == 0).all(), "solver failures in the benchmark batch"
    # ---- multi-rank correctness of the gather: every rank's slice of the root's block == that rank's own result ----
    gather_check = None
    if world > 1:
        own_u, own_x = s.get_u(), s.get_x()
        mine = (float(own_u.sum()), float(np.abs(own_x).sum()), int(st.sum()), own_u[::997].tobytes())
        sums = [None] * world if rank == 0 else None
        dist.gather_object(mine, sums, dst=0)
        if rank == 0:
            ua = np.empty((world * B, N, 2)); xa = np.empty((world * B, N + 1, 7)); sa = np.empty(world * B, dtype=np.int32)
            dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
            _lib.check(L.admpc_batch_get_gathered(s.h, ua.ctypes.data_as(dp), xa.ctypes.data_as(dp), sa.ctypes.data_as(ip)), "get_gathered")
            for r in range(world):
                blk = slice(r * B, (r + 1) * B)
                got = (float(ua[blk].sum()), float(np.abs(xa[blk]).sum()), int(sa[blk].sum()), ua[blk][::997].tobytes())
                assert got == sums[r], "gathered block of rank %d differs from the rank's own result" % r
            gather_check = "root's gathered block == every rank's own (u, x, status): checksums + sampled rows, %d ranks" % world
    n_ipm = float(qi.mean())
    t_local = sum(step_ms)
    if dist is not None:
        import torch
        tt = torch.tensor([t_local], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_all = float(tt.item())
    else:
        t_all = t_local
    ms_per_step = t_all / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # ---- end to end through the public API with pinned HOST buffers ("e2e") --------------------------------
    # PipelinedSolver = the public batched call: 8 chunks on 8 streams so that H2D / solve / D2H overlap.
    from ad_mpc_b200 import PipelinedSolver
    s.set_profiling(False)
    ps = PipelinedSolver(B, opts, device=local, chunks=8)
    ps.set_gp(model)
    pin = {k: PinnedArray(v.shape) for k, v in (("x0", batch["x0"]), ("yref", batch["yref"]))}
    pin["x0"].array[:] = batch["x0"]; pin["yref"].array[:] = batch["yref"]
    pin_p = PinnedArray((B,)); pin_p.array[:] = batch["p"][:, 0]
    out_u, out_x = PinnedArray((B, N, 2)), PinnedArray((B, N + 1, 7))
    out_st = PinnedArray((B,), dtype=np.int32)
    e2e_ms = []
    for it in range(max(args.warmup, 3) + args.steps):
        ps.set_iterate(x_init, u_init)
        ps.wait()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        ps.solve_batch(pin["x0"].array, pin["yref"].array, pin_p.array, out_u.array, out_x.array, out_st.array)
        ms = (time.perf_counter() - t0) * 1e3          # host clock around the blocking public call (4 streams inside)
        if it >= max(args.warmup, 3):
            e2e_ms.append(ms)
    e_local = sum(e2e_ms)
    if dist is not None:
        tt = torch.tensor([e_local], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e_all = float(tt.item())
    else:
        e_all = e_local
    e2e_value = B * world / (e_all / args.steps * 1e-3)
    h2d = int(batch["x0"].nbytes + batch["yref"].nbytes + B * 8)
    d2h = int(B * N * 2 * 8 + B * (N + 1) * 7 * 8 + B * 4)
    assert np.array_equal(out_st.array, st)
    ref_u = s.get_u()
    assert np.array_equal(out_u.array, ref_u), "pipelined e2e result differs from the resident-input result"
    # ---- same call with the reference generated on the device (f1): the host ships vehicle states only -----------
    import math
    Lt = 800
    s_arc = np.arange(Lt) * (2 * math.pi * 50.0 / Lt)
    th = s_arc / 50.0
    track = np.stack([np.full(Lt, 8.0), 50 * np.cos(th), 50 * np.sin(th), (th + math.pi / 2 + math.pi) % (2 * math.pi) - math.pi,
                      s_arc, np.full(Lt, 0.02)], axis=1)
    ps.set_track(track, H=N, traj_dt=CFG["dt"], anchor=True)
    pose_ms = []
    for it in range(3 + args.steps):
        ps.set_iterate(x_init, u_init)
        ps.wait()
        t0 = time.perf_counter()
        ps.solve_from_pose(pin["x0"].array, pin_p.array, out_u.array, out_x.array, out_st.array)
        if it >= 3:
            pose_ms.append((time.perf_counter() - t0) * 1e3)
    e2e_pose = B / (sum(pose_ms) / len(pose_ms) * 1e-3)
    ps.close()

    # ---- the same step on the Frenet-frame model variant (SURVEY 8a A2'; rank 0, reported beside the headline) --------
    frenet = None
    if rank == 0:
        from ad_mpc_b200 import workload as wlf
        from ad_mpc_b200 import kappa_pp_from_knots
        fb = wlf.make_batch_frenet(B, N, seed=20263, p=1.0)

        def frenet_run(opts_f, spline):
            fs = BatchSolver(B, opts_f, device=local)
            fs.set_gp(model)
            fs.set_x0(fb["x0"]); fs.set_yref(fb["yref"]); fs.set_p(fb["p"]); fs.set_kappa(fb["kappa"])
            if spline:      # kappa(s) = 0.02 + 0.01 sin(0.03 s) as the not-a-knot cubic of the reference's bspline interpolant
                kn = np.linspace(-50.0, 450.0, 26)
                bb, cc = kappa_pp_from_knots(kn, 0.02 + 0.01 * np.sin(0.03 * kn))
                fs.set_kappa_spline(np.tile(bb, (B, 1)), np.tile(cc, (B, 1, 1)))
            f_ms = []
            for it in range(3 + args.steps):
                fs.set_iterate(fb["x_init"], fb["u_init"])
                fs.flush_l2()
                fs.timer_start()
                fs.solve()
                ms = fs.timer_stop()
                if it >= 3:
                    f_ms.append(ms)
            fst = fs.get_status()[0]
            fs.close()
            return {"value": B / (statistics.mean(f_ms) * 1e-3), "unit": "solves/s", "ms_per_step": statistics.mean(f_ms),
                    "ok": bool((fst == 0).all())}

        frenet = frenet_run(default_opts(N, model_variant=1), False)
        frenet["note"] = ("Frenet-frame model variant (model_variant=1, curvature per node, shipped constraint set), same batch "
                          "size / horizon / GP, device-resident inputs, one GPU; kernels gp_sweep_kernel<FR> + "
                          "prepare_dense_kernel + qp_mma_g_kernel")
        qf = [0.0, 10.0, 10.0, 10.0, 10.0, 1.0, 0.1]
        own = frenet_run(default_opts(N, model_variant=1, con_set=1, W=qf + [10.0, 10.0], We=[0.01 * v for v in qf],
                                      zl=[100.0, 100.0], zu=[100.0, 100.0], lbu=[-10.0, -2.0], ubu=[5.0, 2.0], lbx=-0.52, ubx=0.52,
                                      lbx2=-2.0, ubx2=2.0), True)
        own["note"] = ("the variant as the reference defines it: kappa(s) spline evaluated inside the model (dense column of s) + "
                       "its own constraint set (con_set=1: acceleration soft, steering rate hard, e_y hard, steering angle soft)")
        frenet["reference_defined"] = own

    # ---- single-instance latency through the acados-shim symbols (cfg 1: nominal, N=20, B=1) ---------------------
    single_ms = []
    if rank == 0:
        from ad_mpc_b200 import AcadosOcpSolverB200
        from ad_mpc_b200 import workload as wl
        b1 = wl.make_batch(1, N, seed=20261, p=0.0)
        cap = AcadosOcpSolverB200(default_opts(N))
        for j in range(N):
            cap.set(j, "yref", b1["yref"][0][j * 9:(j + 1) * 9])
        cap.set(N, "yref", b1["yref"][0][N * 9:])
        for j in range(N + 1):
            cap.set(j, "x", b1["x_init"][0, j])
        x0c = b1["x0"][0].copy()
        for it in range(55):                                   # 5 warm-up + 50 closed-loop RTI steps (SURVEY 8d cfg 1)
            cap.set(0, "lbx", x0c); cap.set(0, "ubx", x0c)
            t0 = time.perf_counter()
            cap.solve()
            single_ms.append((time.perf_counter() - t0) * 1e3)
            x0c = cap.get(1, "x")
        single_ms = sorted(single_ms[5:])

    if rank == 0:
        # ---- roofline of the dominant kernel (FP64 DFMA pipe) ------------------------------------
        peak = C.c_double()
        _lib.check(L.admpc_measure_fp64_peak(local, C.byref(peak)), "fp64 peak")
        ops = algorithmic_ops(N, CFG["M"], CFG["dz"], CFG["n_out"], n_ipm)
        kern = {"prepare": statistics.mean(prep_ms), "qp": statistics.mean(qp_ms)}
        dom = max(kern, key=kern.get)
        achieved = ops[dom] * B / (kern[dom] * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(dom)
            except Exception:
                traffic = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        _v = os.environ.get("ADMPC_QP_VARIANT", "")
        qp_name = "qp_mma_kernel<%d>" % (1 if N <= 31 else 2 if N <= 63 else 4) if _v in ("", "0", "7") and N <= 127 else "qp_warp_kernel"
        roofline = {"bound": "fp64", "kernel": {"prepare": "gp_sweep_kernel + prepare_kernel<GP>", "qp": qp_name}[dom],
                    "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s", "frac": achieved / peak.value,
                    "traffic": traffic,
                    "peak_source": "DFMA microbenchmark admpc_measure_fp64_peak run in this process (MEASURED_PEAKS.json "
                                   "has no FP64 entry; the path is FP64-CUDA-core bound, neither hbm nor tensor)",
                    "flops_per_solve": ops[dom], "n_ipm_mean": n_ipm,
                    "kernels_ms": kern,
                    "kernels_frac": {k: ops[k] * B / (kern[k] * 1e-3) / 1e12 / peak.value for k in kern},
                    "all_kernels_tflops": (ops["prepare"] + ops["qp"]) * B / ((kern["prepare"] + kern["qp"]) * 1e-3) / 1e12,
                    "step_frac": (ops["prepare"] + ops["qp"]) * B / (ms_per_step * 1e-3) / 1e12 / peak.value,     # whole step, all kernels
                    "hbm": {"algorithmic_GBps": algorithmic_bytes(N) * B / (ms_per_step * 1e-3) / 1e9, "peak_GBps": hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}
        # ---- CPU baseline on this box's host cores ----------------------------------------------------------
        cpu = None
        if world == 1:                                                       # reported at N=1 only
            sample = 8192
            rate0, secs0, cores = cpu_oracle_rate(1024)                      # calibrate, then size the sample to ~12 s
            reps = max(1, min(80, int(12.0 * rate0 / sample)))
            rate, secs, cores = cpu_oracle_rate(sample, reps=reps)
            # the other BASELINE.md 2 numbers: single-thread latency on cfg1, all-core throughput on cfg2 and cfg4
            l50, l99 = cpu_single_latency(50)
            r2, s2, _ = cpu_oracle_rate(4096, gp=False, p=0.0, reps=40)
            r4a, _, _ = cpu_oracle_rate(64, N=40, M=2000)                     # calibrate cfg4 (about 20x the work per solve)
            n4 = int(max(64, min(4096, 4.0 * r4a)))
            r4, s4, _ = cpu_oracle_rate(n4, N=40, M=2000)
            cpu = {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port",
                   "sample": "%d instances of the cfg3 workload x%d, OpenMP over instances, %.1f s" % (sample, reps, secs),
                   "build": _ORC_FLAGS,
                   "cfg1_single_thread_latency_ms": {"p50": l50, "p99": l99, "cores": 1,
                                                     "sample": "nominal N=20, B=1, 50 closed-loop RTI steps after 5 warm-up"},
                   "cfg2_nominal_B4096": {"value": r2, "unit": "solves/s", "cores": cores, "sample": "4096 instances x40, %.2f s" % s2},
                   "cfg4_gp2000_N40": {"value": r4, "unit": "solves/s", "cores": cores, "sample": "%d instances, %.2f s" % (n4, s4)},
                   "note": "CPU restatement of the acados path (oracle/); acados itself cannot be built here"}
        e2e_sorted = sorted(e2e_ms)
        line = {"metric": "SQP-RTI MPC solves/sec (N=20, GP-augmented)", "value": value, "unit": "solves/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(config_dict(world), **({"gather": ("fused: QP-kernel epilogue stores into the root's block over NVLink peer memory (CUDA IPC) + 4-byte all-reduce" if gather_mode == 1 else "NCCL grouped send/recv of packed blocks")} if world > 1 else {})),
                "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "p50_ms_per_batch_call": e2e_sorted[len(e2e_sorted) // 2],
                        "p99_ms_per_batch_call": e2e_sorted[min(len(e2e_sorted) - 1, int(0.99 * len(e2e_sorted)))],
                        "pose_only_per_gpu": {"value": e2e_pose, "unit": "solves/s", "h2d_bytes_per_step": B * 8 * 8,
                                              "note": "same call with the reference generated on the device from "
                                                      "vehicle states (refgen_kernel, anchored mode); rank 0"}},
                "gpu_launches": counted, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "host": {"numa_binding": numa, "e2e_entry": "admpc_pipe_solve_host (one C call, 8 chunk streams inside the library)"},
                "gather_check": gather_check,
                "wall_s_timed_region": wall, "frenet_variant": frenet,
                "latency": {"p50_ms_per_batch_solve": sorted(step_ms)[len(step_ms) // 2], "batch": B,
                            "p99_ms_per_batch_solve": sorted(step_ms)[min(len(step_ms) - 1, int(0.99 * len(step_ms)))],
                            "single_instance_p50_ms": single_ms[len(single_ms) // 2] if single_ms else None,
                            "single_instance_p99_ms": single_ms[min(len(single_ms) - 1, int(0.99 * len(single_ms)))] if single_ms else None,
                            "single_instance_note": "cfg1: nominal N=20, B=1, 50 closed-loop RTI steps through "
                                                    "sim_car_acados_solve (host clock around the call: zero-copy pinned I/O block, the step replayed as one CUDA graph)"}}
        print(json.dumps(line), flush=True)
    barrier()
    s.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
