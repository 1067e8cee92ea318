"""ctypes binding of libadmpc_b200.so (include/admpc.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback in the product path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ADMPC_LIB") or os.path.join(HERE, "libadmpc_b200.so")   # ADMPC_LIB: kernel-variant builds (scripts/)

DZMAX, GPOUT_MAX = 8, 2


class AdmpcOpts(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("iter_max", C.c_int), ("gp_enabled", C.c_int), ("gp_nout", C.c_int), ("gp_M", C.c_int),
        ("gp_dz", C.c_int), ("gp_stage0_trigger", C.c_int), ("model_variant", C.c_int),
        ("gp_feat", C.c_int * DZMAX), ("gp_row", C.c_int * GPOUT_MAX), ("gp_precision", C.c_int), ("con_set", C.c_int),
        ("dt", C.c_double), ("W", C.c_double * 9), ("We", C.c_double * 7),
        ("zl", C.c_double * 2), ("zu", C.c_double * 2), ("Zl", C.c_double * 2), ("Zu", C.c_double * 2),
        ("lbu", C.c_double * 2), ("ubu", C.c_double * 2), ("lbx", C.c_double), ("ubx", C.c_double), ("lbx2", C.c_double), ("ubx2", C.c_double),
        ("mass", C.c_double), ("lf", C.c_double), ("lr", C.c_double), ("iz", C.c_double), ("cf2", C.c_double),
        ("cr2", C.c_double),
        ("mu0", C.c_double), ("tol_stat", C.c_double), ("tol_eq", C.c_double), ("tol_ineq", C.c_double),
        ("tol_comp", C.c_double), ("alpha_min", C.c_double), ("lam_min", C.c_double), ("t_min", C.c_double),
        ("thr0", C.c_double), ("reg", C.c_double), ("blend_min", C.c_double), ("blend_max", C.c_double),
    ]


# every symbol include/admpc.h declares: name -> (restype, argtypes)
_dp, _ip, _vp, _cp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, C.c_char_p
_op = C.POINTER(AdmpcOpts)
SYMBOLS = {
    "admpc_default_opts": (None, [_op]),
    "admpc_last_error": (_cp, []),
    "admpc_device_count": (C.c_int, []),
    "sim_car_acados_create_capsule": (_vp, []),
    "sim_car_acados_free_capsule": (C.c_int, [_vp]),
    "sim_car_acados_create": (C.c_int, [_vp]),
    "sim_car_acados_create_with_discretization": (C.c_int, [_vp, C.c_int, _dp]),
    "sim_car_acados_update_time_steps": (C.c_int, [_vp, C.c_int, _dp]),
    "sim_car_acados_update_params": (C.c_int, [_vp, C.c_int, _dp, C.c_int]),
    "sim_car_acados_solve": (C.c_int, [_vp]),
    "sim_car_acados_reset": (C.c_int, [_vp, C.c_int]),
    "sim_car_acados_free": (C.c_int, [_vp]),
    "sim_car_acados_print_stats": (None, [_vp]),
    "sim_car_acados_set_opts": (C.c_int, [_vp, _op]),
    "sim_car_acados_set": (C.c_int, [_vp, C.c_int, _cp, _dp, C.c_int]),
    "sim_car_acados_get": (C.c_int, [_vp, C.c_int, _cp, _dp, C.c_int]),
    "sim_car_acados_get_stat": (C.c_int, [_vp, _cp, _vp]),
    "admpc_batch_create": (C.c_int, [_op, C.c_int, C.c_int, C.POINTER(_vp)]),
    "admpc_batch_free": (C.c_int, [_vp]),
    "admpc_batch_size": (C.c_int, [_vp]),
    "admpc_batch_horizon": (C.c_int, [_vp]),
    "admpc_batch_set_gp": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int]),
    "admpc_batch_set_gp_ensemble": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int]),
    "admpc_batch_select_gp": (C.c_int, [_vp, _dp, _dp]),
    "admpc_batch_set_gp_index": (C.c_int, [_vp, _ip]),
    "admpc_batch_get_gp_index": (C.c_int, [_vp, _ip]),
    "admpc_batch_set_x0": (C.c_int, [_vp, _dp]),
    "admpc_batch_set_yref": (C.c_int, [_vp, _dp]),
    "admpc_batch_set_p": (C.c_int, [_vp, _dp]),
    "admpc_batch_set_p_scalar": (C.c_int, [_vp, _dp]),
    "admpc_batch_set_kappa": (C.c_int, [_vp, _dp]),
    "admpc_batch_set_kappa_spline": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "admpc_batch_set_gp_state": (C.c_int, [_vp, _dp]),
    "admpc_batch_set_iterate": (C.c_int, [_vp, _dp, _dp]),
    "admpc_batch_set_duals": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp]),
    "sim_car_acados_update_qp_solver_cond_N": (C.c_int, [_vp, C.c_int]),
    "admpc_batch_reset": (C.c_int, [_vp]),
    "admpc_batch_solve": (C.c_int, [_vp]),
    "admpc_batch_wait": (C.c_int, [_vp]),
    "admpc_batch_get_u": (C.c_int, [_vp, _dp]),
    "admpc_batch_get_x": (C.c_int, [_vp, _dp]),
    "admpc_batch_get_pi": (C.c_int, [_vp, _dp]),
    "admpc_batch_get_lam": (C.c_int, [_vp, _dp]),
    "admpc_batch_get_t": (C.c_int, [_vp, _dp]),
    "admpc_batch_get_slacks": (C.c_int, [_vp, _dp, _dp]),
    "admpc_batch_get_status": (C.c_int, [_vp, _ip, _ip, _ip]),
    "admpc_batch_get_lin": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp]),
    "admpc_batch_solve_host": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp, _ip]),
    "admpc_batch_solve_host_async": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp, _ip]),
    "admpc_pipe_create": (C.c_int, [_op, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "admpc_pipe_free": (C.c_int, [_vp]),
    "admpc_pipe_chunks": (C.c_int, [_vp]),
    "admpc_pipe_chunk": (_vp, [_vp, C.c_int, _ip, _ip]),
    "admpc_pipe_set_gp": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int]),
    "admpc_pipe_set_iterate": (C.c_int, [_vp, _dp, _dp]),
    "admpc_pipe_set_track": (C.c_int, [_vp, C.c_int, _dp, C.c_int, C.c_double, C.c_int]),
    "admpc_pipe_solve_host": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp, _ip]),
    "admpc_pipe_solve_host_async": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp, _ip]),
    "admpc_pipe_solve_pose": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _ip]),
    "admpc_pipe_wait": (C.c_int, [_vp]),
    "admpc_pipe_kernel_launches": (C.c_longlong, [_vp]),
    "admpc_batch_set_track": (C.c_int, [_vp, C.c_int, _dp, C.c_int, C.c_double]),
    "admpc_batch_set_track_anchor": (C.c_int, [_vp, C.c_int]),
    "admpc_batch_make_yref": (C.c_int, [_vp]),
    "admpc_batch_solve_pose_async": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _ip]),
    "admpc_batch_get_yref": (C.c_int, [_vp, _dp]),
    "admpc_batch_get_waypoint_info": (C.c_int, [_vp, _dp, _dp, _dp, _ip]),
    "admpc_batch_postsolve": (C.c_int, [_vp, C.c_int, C.c_int]),
    "admpc_batch_closed_loop": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp]),
    "admpc_batch_get_loop_info": (C.c_int, [_vp, _ip, _ip, _ip, _dp, _dp]),
    "admpc_batch_set_profiling": (C.c_int, [_vp, C.c_int]),
    "admpc_batch_last_ms": (C.c_int, [_vp, _cp, C.POINTER(C.c_float)]),
    "admpc_batch_kernel_launches": (C.c_longlong, [_vp]),
    "admpc_batch_timer_start": (C.c_int, [_vp]),
    "admpc_batch_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "admpc_batch_flush_l2": (C.c_int, [_vp]),
    "admpc_host_alloc": (_vp, [C.c_ulonglong]),
    "admpc_host_free": (C.c_int, [_vp]),
    "admpc_nccl_unique_id": (C.c_int, [_vp]),
    "admpc_batch_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "admpc_batch_bcast_gp": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int]),
    "admpc_batch_gather": (C.c_int, [_vp, C.c_int, _dp, _dp, _ip]),
    "admpc_batch_get_gathered": (C.c_int, [_vp, _dp, _dp, _ip]),
    "admpc_batch_gather_enable": (C.c_int, [_vp, C.c_int]),
    "admpc_batch_barrier": (C.c_int, [_vp]),
    "admpc_gp_fit": (C.c_int, [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp, _dp, C.POINTER(C.c_float)]),
    "admpc_gp_predict": (C.c_int, [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int, _dp,
                                   _dp, _dp, _dp]),
    "admpc_batch_solve_sqp": (C.c_int, [_vp, C.c_int, _dp, _ip]),
    "admpc_batch_get_sqp_info": (C.c_int, [_vp, _ip, _ip, _dp]),
    "sim_car_acados_set_nlp_solver": (C.c_int, [_vp, _cp, C.c_int, _dp]),
    "admpc_measure_fp64_peak": (C.c_int, [C.c_int, _dp]),
    "admpc_measure_fp64_mix": (C.c_int, [C.c_int, C.c_int, _dp]),
}

_lib = None


class AdmpcError(RuntimeError):
    pass


def load():
    """Load the CUDA library. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise AdmpcError("libadmpc_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(the solver has no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc < 0:
        raise AdmpcError("%s failed (%d): %s" % (what, rc, load().admpc_last_error().decode()))
    return rc
