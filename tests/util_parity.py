"""Shared helpers for the parity tests: oracle <-> CUDA option mirroring and the mixed abs/rel error metric."""
import numpy as np

from oracle import oracle as orc

OPT_FIELDS = ["N", "iter_max", "dt", "W", "We", "zl", "zu", "Zl", "Zu", "lbu", "ubu", "lbx", "ubx", "mass", "lf", "lr",
              "iz", "cf2", "cr2", "mu0", "tol_stat", "tol_eq", "tol_ineq", "tol_comp", "alpha_min", "lam_min", "t_min",
              "thr0", "reg", "con_set", "lbx2", "ubx2"]


def mirror_opts(gpu_opts):
    """Build the oracle's option struct from the product's (field by field; the structs are independent types)."""
    o = orc.default_opts(N=gpu_opts.N)
    for f in OPT_FIELDS:
        v = getattr(gpu_opts, f)
        if hasattr(v, "__len__"):
            for i in range(len(v)):
                getattr(o, f)[i] = v[i]
        else:
            setattr(o, f, v)
    if getattr(gpu_opts, "model_variant", 0) == 1:
        o.model_backend = 2                      # Frenet variant of the oracle
    return o


def mixed_err(a, b):
    """max |a-b| / max(1,|b|)   (SURVEY 7 hard parts: compare per entry with a mixed abs/rel metric)."""
    a, b = np.asarray(a), np.asarray(b)
    return float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max()) if a.size else 0.0


def oracle_batch(o, batch, gp=None, gp_state=None):
    return orc.rti_batch(o, batch["x0"], batch["yref"], batch["p"], batch["x_init"], batch["u_init"], gp=gp, gp_state=gp_state)
