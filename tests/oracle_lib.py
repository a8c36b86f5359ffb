"""ctypes bindings of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product (nmpc_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")


class DdpConfig(C.Structure):
    """Field-for-field the same layout as nmpc_b200_ddp_config (include/nmpc_b200/c_api.h)."""

    _fields_ = [
        ("horizon_steps", C.c_int),
        ("max_iter", C.c_int),
        ("reg_type", C.c_int),
        ("with_input_constraint", C.c_int),
        ("n_alpha", C.c_int),
        ("reserved", C.c_int),
        ("initial_lambda", C.c_double),
        ("initial_dlambda", C.c_double),
        ("lambda_factor", C.c_double),
        ("lambda_min", C.c_double),
        ("lambda_max", C.c_double),
        ("k_rel_norm_thre", C.c_double),
        ("lambda_thre", C.c_double),
        ("cost_update_ratio_thre", C.c_double),
        ("cost_update_thre", C.c_double),
        ("alpha_list", C.c_double * 16),
    ]


class FmpcConfig(C.Structure):
    """Field-for-field the same layout as nmpc_b200_fmpc_config."""

    _fields_ = [
        ("horizon_steps", C.c_int),
        ("max_iter", C.c_int),
        ("check_nan", C.c_int),
        ("init_complementary_variable", C.c_int),
        ("update_barrier_eps", C.c_int),
        ("break_if_llt_fails", C.c_int),
        ("enable_line_search", C.c_int),
        ("merit_const_scale_from_lagrange_multipliers", C.c_int),
        ("kkt_error_thre", C.c_double),
        ("initial_barrier_eps", C.c_double),
    ]


def _host_cpu_tag():
    """Short hash of this host's CPU model and instruction-set flags: -march=native code must not travel."""
    import hashlib

    ident = []
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith(("model name", "flags")):
                    ident.append(line.split(":", 1)[1].strip())
                    if len(ident) == 2:
                        break
    except OSError:
        pass
    return hashlib.sha1("|".join(ident).encode()).hexdigest()[:10]


def build(native=False):
    """Compile the oracle if needed; returns the path of the shared library."""
    target = f"_build/liboracle_native_{_host_cpu_tag()}.so" if native else "_build/liboracle.so"
    subprocess.run(["make", "-s", "-C", _ORACLE_DIR, target], check=True)
    return os.path.join(_ORACLE_DIR, target)


_libs = {}


def lib(native=False):
    key = bool(native)
    if key not in _libs:
        path = build(native)
        L = C.CDLL(path)
        L.oracle_num_threads.restype = C.c_int
        _libs[key] = L
    return _libs[key]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def model_dims(model, native=False):
    nx, nu, ng, npar = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = lib(native).oracle_model_dims(model.encode(), C.byref(nx), C.byref(nu), C.byref(ng), C.byref(npar))
    if rc != 0:
        raise KeyError(model)
    return nx.value, nu.value, ng.value, npar.value


def default_params(model):
    nparams = model_dims(model)[3]
    p = np.zeros(nparams)
    assert lib().oracle_model_default_params(model.encode(), _p(p)) == 0
    return p


def ddp_config(**kw):
    cfg = DdpConfig()
    lib().oracle_ddp_config_default(C.byref(cfg))
    alpha = kw.pop("alpha_list", None)
    if alpha is not None:
        cfg.n_alpha = len(alpha)
        for i, a in enumerate(alpha):
            cfg.alpha_list[i] = a
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def fmpc_config(**kw):
    cfg = FmpcConfig()
    lib().oracle_fmpc_config_default(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def ddp_solve_batch(model, params, cfg, x0, u_init, t0=0.0, u_lo=None, u_hi=None, nthreads=0, native=False,
                    outputs=True):
    """Solve B independent DDP problems with the oracle.  Returns a dict of numpy arrays."""
    nx, nu, _, _ = model_dims(model, native)
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, nx)
    B = x0.shape[0]
    N = cfg.horizon_steps
    u_init = np.ascontiguousarray(u_init, dtype=np.float64).reshape(B, N, nu)
    params = np.ascontiguousarray(params, dtype=np.float64)
    u_lo = None if u_lo is None else np.ascontiguousarray(u_lo, dtype=np.float64)
    u_hi = None if u_hi is None else np.ascontiguousarray(u_hi, dtype=np.float64)
    out = {}
    if outputs:
        out["x"] = np.zeros((B, N + 1, nx))
        out["u"] = np.zeros((B, N, nu))
        out["cost_list"] = np.zeros((B, N + 1))
        out["k"] = np.zeros((B, N, nu))
        out["K"] = np.zeros((B, N, nu * nx))
        out["trace"] = np.zeros((B, cfg.max_iter + 1, 9))
    out["n_trace"] = np.zeros(B, dtype=np.int32)
    out["status"] = np.zeros(B, dtype=np.int32)
    out["iters"] = np.zeros(B, dtype=np.int32)
    out["n_fwd"] = np.zeros(B, dtype=np.int32)
    out["n_bwd"] = np.zeros(B, dtype=np.int32)
    rc = lib(native).oracle_ddp_solve_batch(
        model.encode(), _p(params), C.byref(cfg), C.c_int(B), C.c_double(t0), _p(x0), _p(u_init), _p(u_lo), _p(u_hi),
        _p(out.get("x")), _p(out.get("u")), _p(out.get("cost_list")), _p(out.get("k")), _p(out.get("K")),
        _p(out.get("trace")), _p(out["n_trace"]), _p(out["status"]), _p(out["iters"]), _p(out["n_fwd"]),
        _p(out["n_bwd"]), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle_ddp_solve_batch failed: {rc}")
    if outputs:
        out["cost"] = out["cost_list"].sum(axis=1)
    return out


def fmpc_solve_batch(model, params, cfg, x0, var, t0=0.0, nthreads=0, native=False):
    """var: dict with x[B,N+1,NX], u[B,N,NU], lambda[B,N+1,NX], s[B,N,NG], nu[B,N,NG]."""
    nx, nu, ng, _ = model_dims(model, native)
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, nx)
    B = x0.shape[0]
    N = cfg.horizon_steps
    vin = {k: np.ascontiguousarray(var[k], dtype=np.float64) for k in ("x", "u", "lambda", "s", "nu")}
    assert vin["x"].shape == (B, N + 1, nx) and vin["u"].shape == (B, N, nu)
    assert vin["lambda"].shape == (B, N + 1, nx) and vin["s"].shape == (B, N, ng) and vin["nu"].shape == (B, N, ng)
    params = np.ascontiguousarray(params, dtype=np.float64)
    out = {
        "x": np.zeros((B, N + 1, nx)), "u": np.zeros((B, N, nu)), "lambda": np.zeros((B, N + 1, nx)),
        "s": np.zeros((B, N, ng)), "nu": np.zeros((B, N, ng)), "k": np.zeros((B, N, nu)),
        "K": np.zeros((B, N, nu * nx)), "trace": np.zeros((B, cfg.max_iter, 5)),
        "n_trace": np.zeros(B, dtype=np.int32), "status": np.zeros(B, dtype=np.int32),
    }
    rc = lib(native).oracle_fmpc_solve_batch(
        model.encode(), _p(params), C.byref(cfg), C.c_int(B), C.c_double(t0), _p(x0), _p(vin["x"]), _p(vin["u"]),
        _p(vin["lambda"]), _p(vin["s"]), _p(vin["nu"]), _p(out["x"]), _p(out["u"]), _p(out["lambda"]), _p(out["s"]),
        _p(out["nu"]), _p(out["k"]), _p(out["K"]), _p(out["trace"]), _p(out["n_trace"]), _p(out["status"]),
        C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle_fmpc_solve_batch failed: {rc}")
    return out


def model_eval(model, params, t, x, u):
    nx, nu, ng, _ = model_dims(model)
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    o = {
        "x_next": np.zeros(nx), "costs": np.zeros(2), "Fx": np.zeros(nx * nx), "Fu": np.zeros(nx * nu),
        "Lx": np.zeros(nx), "Lu": np.zeros(nu), "Lxx": np.zeros(nx * nx), "Luu": np.zeros(nu * nu),
        "Lxu": np.zeros(nx * nu), "Vx": np.zeros(nx), "Vxx": np.zeros(nx * nx),
    }
    rc = lib().oracle_model_eval(model.encode(), _p(params), C.c_double(t), _p(x), _p(u), _p(o["x_next"]),
                                 _p(o["costs"]), _p(o["Fx"]), _p(o["Fu"]), _p(o["Lx"]), _p(o["Lu"]), _p(o["Lxx"]),
                                 _p(o["Luu"]), _p(o["Lxu"]), _p(o["Vx"]), _p(o["Vxx"]))
    assert rc == 0
    # column-major -> numpy [row, col]
    o["Fx"] = o["Fx"].reshape(nx, nx).T.copy()
    o["Fu"] = o["Fu"].reshape(nu, nx).T.copy()
    o["Lxx"] = o["Lxx"].reshape(nx, nx).T.copy()
    o["Luu"] = o["Luu"].reshape(nu, nu).T.copy()
    o["Lxu"] = o["Lxu"].reshape(nu, nx).T.copy()
    o["Vxx"] = o["Vxx"].reshape(nx, nx).T.copy()
    o["running_cost"], o["terminal_cost"] = o["costs"]
    if ng > 0:
        g, Cm, Dm = np.zeros(ng), np.zeros(ng * nx), np.zeros(ng * nu)
        assert lib().oracle_ineq_eval(model.encode(), _p(params), C.c_double(t), _p(x), _p(u), _p(g), _p(Cm),
                                      _p(Dm)) == 0
        o["g"], o["C"], o["D"] = g, Cm.reshape(nx, ng).T.copy(), Dm.reshape(nu, ng).T.copy()
    return o


def boxqp_solve(H, g, lower, upper, x0=None):
    H = np.asarray(H, dtype=np.float64)
    n = H.shape[0]
    Hc = np.asfortranarray(H).ravel(order="F").copy()
    g = np.ascontiguousarray(g, dtype=np.float64)
    lower = np.ascontiguousarray(lower, dtype=np.float64)
    upper = np.ascontiguousarray(upper, dtype=np.float64)
    x0 = np.zeros(n) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
    x = np.zeros(n)
    retval, iters = C.c_int(), C.c_int()
    rc = lib().oracle_boxqp_solve(C.c_int(n), _p(Hc), _p(g), _p(lower), _p(upper), _p(x0), _p(x), C.byref(retval),
                                  C.byref(iters))
    assert rc == 0
    return x, retval.value, iters.value


def cartpole_x0(B, seed):
    """Synthetic initial states of SURVEY.md 8(d): columns drawn in order pos, theta, vel, omega."""
    rng = np.random.default_rng(seed)
    x0 = np.empty((B, 4))
    x0[:, 0] = rng.uniform(-2.0, 2.0, B)
    x0[:, 1] = rng.uniform(-np.pi, np.pi, B)
    x0[:, 2] = rng.uniform(-1.0, 1.0, B)
    x0[:, 3] = rng.uniform(-1.0, 1.0, B)
    return x0


def ddp_solve_cartpole_tv_limits(params, cfg, x0, u_init, u_lo_steps, u_hi_steps, t0=0.0):
    """Cart-pole DDP with input limits that change along the horizon: u_lo_steps / u_hi_steps [N, 1]."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, 4)
    B, N = x0.shape[0], cfg.horizon_steps
    u_init = np.ascontiguousarray(u_init, dtype=np.float64).reshape(B, N, 1)
    lo = np.ascontiguousarray(u_lo_steps, dtype=np.float64).reshape(N, 1)
    hi = np.ascontiguousarray(u_hi_steps, dtype=np.float64).reshape(N, 1)
    params = np.ascontiguousarray(params, dtype=np.float64)
    out = {"x": np.zeros((B, N + 1, 4)), "u": np.zeros((B, N, 1)), "cost_list": np.zeros((B, N + 1)),
           "status": np.zeros(B, dtype=np.int32), "iters": np.zeros(B, dtype=np.int32)}
    rc = lib().oracle_ddp_solve_batch_cartpole_tv(_p(params), C.byref(cfg), C.c_int(B), C.c_double(t0), _p(x0), _p(u_init),
                                                  _p(lo), _p(hi), _p(out["x"]), _p(out["u"]), _p(out["cost_list"]),
                                                  _p(out["status"]), _p(out["iters"]), C.c_int(0))
    if rc != 0:
        raise RuntimeError(f"oracle_ddp_solve_batch_cartpole_tv failed: {rc}")
    out["cost"] = out["cost_list"].sum(axis=1)
    return out
