"""ctypes binding of include/nmpc_b200/c_api.h (libnmpc_b200.so, built in-tree by __graft_entry__.build())."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NMPC_B200_LIBRARY: load another build of the same library (kernel-variant experiments); the default is the in-tree one
_LIB_PATH = os.environ.get("NMPC_B200_LIBRARY") or os.path.join(_HERE, "libnmpc_b200.so")

OK = 0
ERR_INVALID_ARGUMENT = 1
ERR_RUNTIME = 2
ERR_UNKNOWN_MODEL = 3
ERR_NO_DEVICE = 4
ERR_CUDA = 5
ERR_CAPACITY = 6
ERR_UNSUPPORTED = 7


class NmpcB200Error(RuntimeError):
    """Raised for every non-zero status of the C ABI; ``code`` is the nmpc_b200_status."""

    def __init__(self, code, message):
        super().__init__(f"[nmpc_b200 status {code}] {message}")
        self.code = code
        self.message = message


class InvalidArgument(NmpcB200Error, ValueError):
    """std::invalid_argument of the reference (e.g. DDPSolver.hpp:41-45)."""


class DdpConfigStruct(C.Structure):
    _fields_ = [
        ("horizon_steps", C.c_int),
        ("max_iter", C.c_int),
        ("reg_type", C.c_int),
        ("with_input_constraint", C.c_int),
        ("n_alpha", C.c_int),
        ("use_state_eq_second_derivative", C.c_int),
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


class MpcConfigStruct(C.Structure):
    """nmpc_b200_mpc_config (c_api.h): the reference's host MPC loops, run on the device."""
    _fields_ = [
        ("n_ticks", C.c_int),
        ("plant", C.c_int),
        ("shift_inputs", C.c_int),
        ("clamp_u0", C.c_int),
        ("n_substeps", C.c_int),
        ("feedback", C.c_int),
        ("tick_dt", C.c_double),
        ("sim_dt", C.c_double),
    ]


class FmpcConfigStruct(C.Structure):
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


# every symbol include/nmpc_b200/c_api.h declares (checked by tests/test_capi_symbols.py)
EXPORTED_SYMBOLS = [
    "nmpc_b200_last_error", "nmpc_b200_version", "nmpc_b200_device_count", "nmpc_b200_model_dims",
    "nmpc_b200_model_default_params", "nmpc_b200_model_count", "nmpc_b200_model_name", "nmpc_b200_model_eval",
    "nmpc_b200_ddp_config_default", "nmpc_b200_ddp_create", "nmpc_b200_ddp_destroy", "nmpc_b200_ddp_set_config",
    "nmpc_b200_ddp_get_config", "nmpc_b200_ddp_set_input_limits", "nmpc_b200_ddp_set_input_limits_horizon",
    "nmpc_b200_ddp_solve", "nmpc_b200_ddp_get",
    "nmpc_b200_ddp_sync", "nmpc_b200_ddp_enable_timing", "nmpc_b200_ddp_get_durations", "nmpc_b200_ddp_run_mpc",
    "nmpc_b200_ddp_get_iteration_durations", "nmpc_b200_load_plugin", "nmpc_b200_ddp_set_input_limits_mpc",
    "nmpc_b200_fmpc_config_default", "nmpc_b200_fmpc_create", "nmpc_b200_fmpc_destroy", "nmpc_b200_fmpc_set_config",
    "nmpc_b200_fmpc_solve", "nmpc_b200_fmpc_get", "nmpc_b200_fmpc_sync", "nmpc_b200_fmpc_enable_timing",
    "nmpc_b200_fmpc_get_durations", "nmpc_b200_fmpc_run_mpc",
    "nmpc_b200_ddp_set_tuning", "nmpc_b200_ddp_get_tuning",
    "nmpc_b200_ddp_create_sharded", "nmpc_b200_ddp_sharded_destroy", "nmpc_b200_ddp_sharded_num_shards",
    "nmpc_b200_ddp_sharded_shard", "nmpc_b200_ddp_sharded_range", "nmpc_b200_ddp_sharded_set_config",
    "nmpc_b200_ddp_sharded_set_input_limits", "nmpc_b200_ddp_sharded_solve", "nmpc_b200_ddp_sharded_get",
    "nmpc_b200_fmpc_create_sharded", "nmpc_b200_fmpc_sharded_destroy", "nmpc_b200_fmpc_sharded_num_shards",
    "nmpc_b200_fmpc_sharded_set_config", "nmpc_b200_fmpc_sharded_solve", "nmpc_b200_fmpc_sharded_get",
    "nmpc_b200_peer_buffer_create", "nmpc_b200_peer_buffer_open", "nmpc_b200_peer_buffer_close",
    "nmpc_b200_peer_buffer_destroy", "nmpc_b200_peer_signal", "nmpc_b200_peer_wait", "nmpc_b200_peer_check",
]

_lib = None


def library_path():
    return _LIB_PATH


def lib():
    """Load libnmpc_b200.so (fails loudly when the CUDA extension has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(
                f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nmpc_b200 has no CPU fallback)")
        L = C.CDLL(_LIB_PATH, mode=C.RTLD_GLOBAL)
        L.nmpc_b200_last_error.restype = C.c_char_p
        L.nmpc_b200_model_name.restype = C.c_char_p
        L.nmpc_b200_ddp_config_default.restype = None
        L.nmpc_b200_fmpc_config_default.restype = None
        L.nmpc_b200_ddp_solve.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          C.c_void_p]
        L.nmpc_b200_ddp_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.nmpc_b200_ddp_run_mpc.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.nmpc_b200_ddp_create.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.nmpc_b200_ddp_destroy.argtypes = [C.c_void_p]
        L.nmpc_b200_ddp_set_config.argtypes = [C.c_void_p, C.c_void_p]
        L.nmpc_b200_ddp_get_config.argtypes = [C.c_void_p, C.c_void_p]
        L.nmpc_b200_ddp_set_input_limits.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.nmpc_b200_ddp_set_input_limits_horizon.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.nmpc_b200_ddp_sync.argtypes = [C.c_void_p]
        L.nmpc_b200_ddp_enable_timing.argtypes = [C.c_void_p, C.c_int]
        L.nmpc_b200_ddp_get_durations.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.nmpc_b200_load_plugin.argtypes = [C.c_char_p]
        L.nmpc_b200_ddp_set_input_limits_mpc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.nmpc_b200_ddp_get_iteration_durations.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.nmpc_b200_ddp_set_tuning.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.nmpc_b200_ddp_get_tuning.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.nmpc_b200_ddp_create_sharded.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                   C.c_int, C.c_void_p]
        L.nmpc_b200_ddp_sharded_destroy.argtypes = [C.c_void_p]
        L.nmpc_b200_ddp_sharded_num_shards.argtypes = [C.c_void_p]
        L.nmpc_b200_ddp_sharded_shard.argtypes = [C.c_void_p, C.c_int]
        L.nmpc_b200_ddp_sharded_shard.restype = C.c_void_p
        L.nmpc_b200_ddp_sharded_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nmpc_b200_ddp_sharded_set_config.argtypes = [C.c_void_p, C.c_void_p]
        L.nmpc_b200_ddp_sharded_set_input_limits.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.nmpc_b200_ddp_sharded_solve.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        L.nmpc_b200_ddp_sharded_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
        L.nmpc_b200_fmpc_create_sharded.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                    C.c_int, C.c_void_p]
        L.nmpc_b200_fmpc_sharded_destroy.argtypes = [C.c_void_p]
        L.nmpc_b200_fmpc_sharded_num_shards.argtypes = [C.c_void_p]
        L.nmpc_b200_fmpc_sharded_set_config.argtypes = [C.c_void_p, C.c_void_p]
        L.nmpc_b200_fmpc_sharded_solve.argtypes = [C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 6 + [C.c_int]
        L.nmpc_b200_fmpc_sharded_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
        L.nmpc_b200_peer_buffer_create.argtypes = [C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
        L.nmpc_b200_peer_buffer_open.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.nmpc_b200_peer_buffer_close.argtypes = [C.c_void_p, C.c_int]
        L.nmpc_b200_peer_buffer_destroy.argtypes = [C.c_void_p, C.c_int]
        L.nmpc_b200_peer_signal.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, C.c_void_p]
        L.nmpc_b200_peer_wait.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong, C.c_int, C.c_int, C.c_void_p]
        L.nmpc_b200_peer_check.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.nmpc_b200_fmpc_solve.argtypes = [C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 6 + [C.c_int, C.c_int,
                                                                                                  C.c_void_p]
        L.nmpc_b200_fmpc_run_mpc.argtypes = [C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 6 + [C.c_int] + [
            C.c_void_p] * 5 + [C.c_int, C.c_void_p]
        L.nmpc_b200_fmpc_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.nmpc_b200_fmpc_create.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.nmpc_b200_fmpc_destroy.argtypes = [C.c_void_p]
        L.nmpc_b200_fmpc_set_config.argtypes = [C.c_void_p, C.c_void_p]
        L.nmpc_b200_fmpc_sync.argtypes = [C.c_void_p]
        L.nmpc_b200_fmpc_enable_timing.argtypes = [C.c_void_p, C.c_int]
        L.nmpc_b200_fmpc_get_durations.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(status):
    if status != OK:
        msg = lib().nmpc_b200_last_error().decode(errors="replace")
        if status == ERR_INVALID_ARGUMENT:
            raise InvalidArgument(status, msg)
        raise NmpcB200Error(status, msg)


def device_count():
    return int(lib().nmpc_b200_device_count())


def load_plugin(path):
    """nmpc_b200_load_plugin: a shared library with user problem functors (include/nmpc_b200/plugin.h)."""
    check(lib().nmpc_b200_load_plugin(os.fsencode(path)))


def model_names():
    L = lib()
    return [L.nmpc_b200_model_name(i).decode() for i in range(L.nmpc_b200_model_count())]


def model_dims(model):
    nx, nu, ng, npar = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(lib().nmpc_b200_model_dims(model.encode(), C.byref(nx), C.byref(nu), C.byref(ng), C.byref(npar)))
    return nx.value, nu.value, ng.value, npar.value


def model_default_params(model):
    npar = model_dims(model)[3]
    p = np.zeros(npar)
    check(lib().nmpc_b200_model_default_params(model.encode(), p.ctypes.data_as(C.c_void_p)))
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def model_eval(model, t, x, u, params=None, device=0):
    """Evaluate a problem functor on the GPU at n points; returns a dict of numpy arrays ([n, rows, cols])."""
    nx, nu, ng, npar = model_dims(model)
    params = model_default_params(model) if params is None else np.ascontiguousarray(params, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, nx)
    n = x.shape[0]
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(n, nu)
    t = np.ascontiguousarray(np.broadcast_to(np.asarray(t, dtype=np.float64), (n,)))
    o = {
        "x_next": np.zeros((n, nx)), "running_cost": np.zeros(n), "terminal_cost": np.zeros(n),
        "Fx": np.zeros((n, nx * nx)), "Fu": np.zeros((n, nx * nu)), "Lx": np.zeros((n, nx)), "Lu": np.zeros((n, nu)),
        "Lxx": np.zeros((n, nx * nx)), "Luu": np.zeros((n, nu * nu)), "Lxu": np.zeros((n, nx * nu)),
        "Vx": np.zeros((n, nx)), "Vxx": np.zeros((n, nx * nx)),
        "g": np.zeros((n, max(ng, 1))), "C": np.zeros((n, max(ng * nx, 1))), "D": np.zeros((n, max(ng * nu, 1))),
    }
    check(lib().nmpc_b200_model_eval(
        model.encode(), _ptr(params), C.c_int(params.size), C.c_int(device), C.c_int(n), _ptr(t), _ptr(x), _ptr(u),
        _ptr(o["x_next"]), _ptr(o["running_cost"]), _ptr(o["terminal_cost"]), _ptr(o["Fx"]), _ptr(o["Fu"]),
        _ptr(o["Lx"]), _ptr(o["Lu"]), _ptr(o["Lxx"]), _ptr(o["Luu"]), _ptr(o["Lxu"]), _ptr(o["Vx"]), _ptr(o["Vxx"]),
        _ptr(o["g"]), _ptr(o["C"]), _ptr(o["D"])))

    def colmajor(a, r, c):
        return a.reshape(n, c, r).transpose(0, 2, 1).copy()

    o["Fx"] = colmajor(o["Fx"], nx, nx)
    o["Fu"] = colmajor(o["Fu"], nx, nu)
    o["Lxx"] = colmajor(o["Lxx"], nx, nx)
    o["Luu"] = colmajor(o["Luu"], nu, nu)
    o["Lxu"] = colmajor(o["Lxu"], nx, nu)
    o["Vxx"] = colmajor(o["Vxx"], nx, nx)
    if ng > 0:
        o["g"] = o["g"][:, :ng]
        o["C"] = colmajor(o["C"], ng, nx)
        o["D"] = colmajor(o["D"], ng, nu)
    else:
        del o["g"], o["C"], o["D"]
    return o


def as_device_or_host(a, shape=None):
    """Classify an array argument: returns (pointer, on_device, keepalive).

    numpy arrays (or anything numpy can convert) are host buffers; torch CUDA tensors are device
    buffers and are passed by pointer without a copy."""
    try:
        import torch
    except ImportError:  # pragma: no cover
        torch = None
    if torch is not None and isinstance(a, torch.Tensor):
        if a.dtype != torch.float64:
            raise InvalidArgument(ERR_INVALID_ARGUMENT, "tensors must be float64")
        a = a.contiguous()
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise InvalidArgument(ERR_INVALID_ARGUMENT, f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
        if a.is_cuda:
            return C.c_void_p(a.data_ptr()), True, a
        a = a.numpy()
    arr = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        if arr.size != int(np.prod(shape)):
            raise InvalidArgument(ERR_INVALID_ARGUMENT, f"expected shape {tuple(shape)}, got {arr.shape}")
        arr = arr.reshape(shape)
    return arr.ctypes.data_as(C.c_void_p), False, arr
