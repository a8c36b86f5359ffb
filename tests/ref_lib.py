"""ctypes binding of oracle/_ref/libnmpc_ref.so: the REFERENCE's own DDPSolver / BoxQP / FmpcSolver headers
compiled (unmodified, from /root/reference) against the Eigen-subset shim of oracle/ref/eigen_shim.

TEST INFRASTRUCTURE.  Exists only where /root/reference does (the build container); everywhere else the
vectors it produced are read from tests/golden/reference_outputs.npz."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_lib as O

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF_DIR = os.path.join(_ROOT, "oracle", "ref")
_LIB = os.path.join(_ROOT, "oracle", "_ref", "libnmpc_ref.so")
REFERENCE_ROOT = os.environ.get("REF", "/root/reference")  # the same variable oracle/ref/Makefile reads


def available():
    return os.path.isdir(REFERENCE_ROOT)


_lib = None


def lib():
    """oracle/_ref/libnmpc_ref.so (reference headers + Eigen shim); NMPC_REF_EIGEN_LIB selects a build of the same sources
    against a real Eigen instead (oracle/ref/Makefile target `eigen`)."""
    global _lib
    if _lib is None:
        override = os.environ.get("NMPC_REF_EIGEN_LIB")
        if override:
            _lib = C.CDLL(override)
            return _lib
        if not available():
            raise RuntimeError("the reference checkout is not present on this machine")
        subprocess.run(["make", "-s", "-C", _REF_DIR, f"REF={REFERENCE_ROOT}"], check=True)
        _lib = C.CDLL(_LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def ddp_config(**kw):
    """The reference's own default Configuration (built by its constructor), then overrides."""
    cfg = O.DdpConfig()
    lib().ref_ddp_config_default(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def ddp_solve_cartpole(params, cfg, x0, u_init, t0=0.0, u_lo=None, u_hi=None):
    N = cfg.horizon_steps
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(4)
    u_init = np.ascontiguousarray(u_init, dtype=np.float64).reshape(N)
    params = np.ascontiguousarray(params, dtype=np.float64)
    out = {"x": np.zeros((N + 1, 4)), "u": np.zeros((N, 1)), "cost_list": np.zeros(N + 1),
           "trace": np.zeros((cfg.max_iter + 1, 9))}
    n_trace, ret = C.c_int(), C.c_int()
    lo = None if u_lo is None else np.ascontiguousarray(u_lo, dtype=np.float64)
    hi = None if u_hi is None else np.ascontiguousarray(u_hi, dtype=np.float64)
    rc = lib().ref_ddp_solve_cartpole(_p(params), C.byref(cfg), C.c_double(t0), _p(x0), _p(u_init), _p(lo), _p(hi),
                                      _p(out["x"]), _p(out["u"]), _p(out["cost_list"]), _p(out["trace"]),
                                      C.byref(n_trace), C.byref(ret))
    if rc != 0:
        raise RuntimeError("reference DDPSolver::solve threw")
    out["n_trace"], out["solve_ret"] = n_trace.value, ret.value
    return out


def ddp_solve_planar(params, cfg, x0, u_init, t0=0.0, u_lo=None, u_hi=None):
    """The reference's DDPSolver<6, 2> on the planar quadrotor (oracle/ref/ref_models.h); BoxQP<2> when limits are on."""
    return _ddp_solve_nxnu("ref_ddp_solve_planar", 6, 2, params, cfg, x0, u_init, t0, u_lo, u_hi)


def ddp_solve_quadrotor(params, cfg, x0, u_init, t0=0.0, u_lo=None, u_hi=None):
    """The reference's DDPSolver<12, 4> on the 3-D quadrotor functor (include/nmpc_b200/models/quadrotor.h behind the
    reference's DDPProblem interface); BoxQP<4> when limits are on."""
    return _ddp_solve_nxnu("ref_ddp_solve_quadrotor", 12, 4, params, cfg, x0, u_init, t0, u_lo, u_hi)


def _ddp_solve_nxnu(entry, nx, nu, params, cfg, x0, u_init, t0, u_lo, u_hi):
    N = cfg.horizon_steps
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(nx)
    u_init = np.ascontiguousarray(u_init, dtype=np.float64).reshape(N, nu)
    params = np.ascontiguousarray(params, dtype=np.float64)
    out = {"x": np.zeros((N + 1, nx)), "u": np.zeros((N, nu)), "cost_list": np.zeros(N + 1),
           "trace": np.zeros((cfg.max_iter + 1, 9))}
    n_trace, ret = C.c_int(), C.c_int()
    lo = None if u_lo is None else np.ascontiguousarray(u_lo, dtype=np.float64).reshape(nu)
    hi = None if u_hi is None else np.ascontiguousarray(u_hi, dtype=np.float64).reshape(nu)
    rc = getattr(lib(), entry)(_p(params), C.byref(cfg), C.c_double(t0), _p(x0), _p(u_init), _p(lo), _p(hi),
                                    _p(out["x"]), _p(out["u"]), _p(out["cost_list"]), _p(out["trace"]),
                                    C.byref(n_trace), C.byref(ret))
    if rc != 0:
        raise RuntimeError("reference DDPSolver::solve threw")
    out["n_trace"], out["solve_ret"] = n_trace.value, ret.value
    return out


def fmpc_solve(model, params, cfg, x0, var, t0=0.0):
    nx, nu, ng, _ = O.model_dims(model)
    N = cfg.horizon_steps
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(nx)
    vin = {k: np.ascontiguousarray(var[k], dtype=np.float64) for k in ("x", "u", "lambda", "s", "nu")}
    out = {"x": np.zeros((N + 1, nx)), "u": np.zeros((N, nu)), "lambda": np.zeros((N + 1, nx)),
           "s": np.zeros((N, ng)), "nu": np.zeros((N, ng)), "K": np.zeros((N, nu * nx)),
           "kkt": np.zeros(max(cfg.max_iter, 1))}
    n_trace, status = C.c_int(), C.c_int()
    params = np.ascontiguousarray(params, dtype=np.float64)
    rc = lib().ref_fmpc_solve(model.encode(), _p(params), C.byref(cfg), C.c_double(t0), _p(x0), _p(vin["x"]),
                              _p(vin["u"]), _p(vin["lambda"]), _p(vin["s"]), _p(vin["nu"]), _p(out["x"]),
                              _p(out["u"]), _p(out["lambda"]), _p(out["s"]), _p(out["nu"]), _p(out["K"]),
                              _p(out["kkt"]), C.byref(n_trace), C.byref(status))
    if rc != 0:
        raise RuntimeError("reference FmpcSolver::solve threw")
    out["n_trace"], out["status"] = n_trace.value, status.value
    return out


def boxqp_solve2(H, g, lower, upper, dynamic):
    Hc = np.asfortranarray(np.asarray(H, dtype=np.float64)).ravel(order="F").copy()
    g, lower, upper = (np.ascontiguousarray(v, dtype=np.float64) for v in (g, lower, upper))
    x = np.zeros(2)
    retval = C.c_int()
    assert lib().ref_boxqp_solve2(_p(Hc), _p(g), _p(lower), _p(upper), C.c_int(int(dynamic)), _p(x),
                                  C.byref(retval)) == 0
    return x, retval.value


def vertical_mpc(horizon_steps, with_constraint, n_ticks, x0=(1.2, 0.0), t0=0.0):
    """TestDDPVerticalMotion's MPC loop (TestDDPVerticalMotion.cpp:236-330) with the reference's
    DDPSolver<2, Eigen::Dynamic>: per-tick current_x, u_list[0] (padded to 2), its size, iterations; and the full
    trajectories of the last solve (u padded to 2)."""
    N, T = int(horizon_steps), int(n_ticks)
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(2)
    out = {"x_log": np.zeros((T, 2)), "u0_log": np.zeros((T, 2)), "dim_log": np.zeros(T, dtype=np.int32),
           "iters_log": np.zeros(T, dtype=np.int32), "x": np.zeros((N + 1, 2)), "u": np.zeros((N, 2))}
    rc = lib().ref_vertical_mpc(N, int(with_constraint), T, C.c_double(t0), _p(x0), _p(out["x_log"]), _p(out["u0_log"]),
                                _p(out["dim_log"]), _p(out["iters_log"]), _p(out["x"]), _p(out["u"]))
    if rc != 0:
        raise RuntimeError("reference DDPSolver<2, Dynamic>::solve threw")
    return out


def centroidal_mpc(horizon_steps, n_ticks, first_max_iter=500):
    """TestDDPCentroidalMotion's MPC loop (TestDDPCentroidalMotion.cpp:238-353) with the reference's
    DDPSolver<9, Eigen::Dynamic>: per-tick current_x, u_list[0] (padded to 16), its size, iterations; the full
    trajectories and cost of the first solve and the trajectories of the last one (u padded to 16)."""
    N, T = int(horizon_steps), int(n_ticks)
    out = {"x_log": np.zeros((T, 9)), "u0_log": np.zeros((T, 16)), "dim_log": np.zeros(T, dtype=np.int32),
           "iters_log": np.zeros(T, dtype=np.int32), "x_first": np.zeros((N + 1, 9)), "u_first": np.zeros((N, 16)),
           "cost_first": np.zeros(1), "x": np.zeros((N + 1, 9)), "u": np.zeros((N, 16))}
    rc = lib().ref_centroidal_mpc(N, int(first_max_iter), T, _p(out["x_log"]), _p(out["u0_log"]), _p(out["dim_log"]),
                                  _p(out["iters_log"]), _p(out["x_first"]), _p(out["u_first"]), _p(out["cost_first"]),
                                  _p(out["x"]), _p(out["u"]))
    if rc != 0:
        raise RuntimeError("reference DDPSolver<9, Dynamic>::solve threw")
    return out
