"""Python mirror of ``nmpc_ddp::DDPSolver`` for batches of independent instances on one GPU.

Names and semantics follow the reference (isri-aist/NMPC nmpc_ddp/include/nmpc_ddp/DDPSolver.h):
``DDPSolver(problem)``, ``config()``, ``solve(current_t, current_x, initial_u_list)``,
``setInputLimitsFunc`` (constant limits), ``controlData()``, ``traceDataList()``,
``computationDuration()``, ``dumpTraceDataList(path)``; the batched entry point is ``solve_batch``.
All compute happens in libnmpc_b200.so; this file only marshals arrays.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _capi
from ._capi import DdpConfigStruct, InvalidArgument, check, lib

# nmpc_b200_ddp_field
F_X, F_U, F_COST_LIST, F_K_FF, F_K_FB, F_TRACE, F_STATUS, F_ITERS, F_N_FORWARD, F_N_BACKWARD, F_COST, F_U0, F_N_TRACE = \
    range(13)

TRACE_FIELDS = ("iter", "cost", "lambda", "dlambda", "alpha", "k_rel_norm", "cost_update_actual",
                "cost_update_expected", "cost_update_ratio")


def _default_alpha_list():
    # DDPSolver.h:53-59: 10^LinSpaced(11, 0, -3)
    return [10.0 ** (0.0 + i * (-3.0 / 10)) if i < 10 else 10.0 ** -3.0 for i in range(11)]


@dataclass
class DDPConfiguration:
    """DDPSolver::Configuration (DDPSolver.h:47-110), same names and defaults."""
    print_level: int = 1
    use_state_eq_second_derivative: bool = False
    with_input_constraint: bool = False
    max_iter: int = 500
    horizon_steps: int = 100
    reg_type: int = 1
    initial_lambda: float = 1e-4
    initial_dlambda: float = 1.0
    lambda_factor: float = 1.6
    lambda_min: float = 1e-6
    lambda_max: float = 1e10
    k_rel_norm_thre: float = 1e-4
    lambda_thre: float = 1e-5
    alpha_list: List[float] = field(default_factory=_default_alpha_list)
    cost_update_ratio_thre: float = 0.0
    cost_update_thre: float = 1e-7

    def to_struct(self):
        s = DdpConfigStruct()
        s.horizon_steps = int(self.horizon_steps)
        s.max_iter = int(self.max_iter)
        s.reg_type = int(self.reg_type)
        s.with_input_constraint = int(bool(self.with_input_constraint))
        if len(self.alpha_list) > 16:
            raise InvalidArgument(_capi.ERR_INVALID_ARGUMENT, "alpha_list longer than 16")
        s.n_alpha = len(self.alpha_list)
        s.use_state_eq_second_derivative = int(bool(self.use_state_eq_second_derivative))
        for name in ("initial_lambda", "initial_dlambda", "lambda_factor", "lambda_min", "lambda_max",
                     "k_rel_norm_thre", "lambda_thre", "cost_update_ratio_thre", "cost_update_thre"):
            setattr(s, name, float(getattr(self, name)))
        for i, a in enumerate(self.alpha_list):
            s.alpha_list[i] = float(a)
        return s


@dataclass
class ControlData:
    """DDPSolver::ControlData (DDPSolver.h:113-123) for the whole batch."""
    x_list: np.ndarray  # [B, N+1, NX]
    u_list: np.ndarray  # [B, N, NU]
    cost_list: np.ndarray  # [B, N+1]


@dataclass
class TraceData:
    """DDPSolver::TraceData (DDPSolver.h:179-216) of one instance and iteration."""
    iter: int = 0
    cost: float = 0.0
    lambda_: float = 0.0
    dlambda: float = 0.0
    alpha: float = 0.0
    k_rel_norm: float = 0.0
    cost_update_actual: float = 0.0
    cost_update_expected: float = 0.0
    cost_update_ratio: float = 0.0
    duration_derivative: float = 0.0
    duration_backward: float = 0.0
    duration_forward: float = 0.0


class DDPSolver:
    """A batch of ``nmpc_ddp::DDPSolver`` objects sharing one problem functor, resident on one B200."""

    def __init__(self, problem, params=None, batch_capacity=1, device=0, config=None):
        """``problem``: name of a registered problem functor (e.g. "cartpole"); ``params``: its flat
        parameter vector (defaults to the functor's defaults)."""
        self._h = C.c_void_p()
        self.problem = problem
        self.nx, self.nu, self.ng, self.n_params = _capi.model_dims(problem)
        self.params = (_capi.model_default_params(problem) if params is None else np.ascontiguousarray(
            params, dtype=np.float64))
        self.device = device
        self.batch_capacity = int(batch_capacity)
        self._config = config if config is not None else DDPConfiguration()
        self._applied = None
        self._B = 0
        self._timing = False
        self._limits_func = None
        s = self._config.to_struct()
        check(lib().nmpc_b200_ddp_create(problem.encode(), self.params.ctypes.data_as(C.c_void_p),
                                         int(self.params.size), C.byref(s), self.batch_capacity, int(device),
                                         C.byref(self._h)))
        self._applied = bytes(s)

    # -- reference API ------------------------------------------------------------------------
    def config(self):
        """Accessor to the configuration (DDPSolver.h:258-267); changes take effect at the next solve."""
        return self._config

    def setInputLimitsFunc(self, input_limits):
        """DDPSolver::setInputLimitsFunc (DDPSolver.h:282-285).  ``input_limits``: a callable t -> (lower, upper), each of
        length NU, evaluated at every horizon time ``current_t + i * dt`` of each solve like the reference's
        input_limits_func_ (DDPSolver.hpp:470); or a constant (lower, upper) pair."""
        if callable(input_limits):
            self._limits_func = input_limits
            return
        self._limits_func = None
        self._apply_config()
        lo = np.ascontiguousarray(input_limits[0], dtype=np.float64).reshape(self.nu)
        hi = np.ascontiguousarray(input_limits[1], dtype=np.float64).reshape(self.nu)
        check(lib().nmpc_b200_ddp_set_input_limits(self._h, lo.ctypes.data_as(C.c_void_p),
                                                   hi.ctypes.data_as(C.c_void_p)))

    def _apply_limits(self, current_t):
        """Evaluate a limits callable over the horizon of the coming solve."""
        func = getattr(self, "_limits_func", None)
        if func is None:
            return
        N = self._config.horizon_steps
        dt = float(self.params[0])  # every functor of this library keeps dt first in its flat parameter vector
        lo, hi = np.empty((N, self.nu)), np.empty((N, self.nu))
        for i in range(N):
            l, h = func(current_t + i * dt)
            lo[i], hi[i] = np.asarray(l, dtype=np.float64).reshape(self.nu), np.asarray(h, dtype=np.float64).reshape(self.nu)
        check(lib().nmpc_b200_ddp_set_input_limits_horizon(self._h, N, lo.ctypes.data_as(C.c_void_p),
                                                           hi.ctypes.data_as(C.c_void_p)))

    def _apply_limits_mpc(self, current_t, n_ticks, tick_dt):
        """A limits callable at every (tick, horizon step) time of the device-resident MPC loop (the reference
        evaluates input_limits_func_ at every solve, DDPSolver.hpp:470); uploaded only when it depends on time."""
        func = getattr(self, "_limits_func", None)
        if func is None or n_ticks <= 0:
            return
        N = self._config.horizon_steps
        dt = float(self.params[0])
        lo, hi = np.empty((n_ticks, N, self.nu)), np.empty((n_ticks, N, self.nu))
        for k in range(n_ticks):
            for i in range(N):
                l, h = func(current_t + k * tick_dt + i * dt)
                lo[k, i], hi[k, i] = np.asarray(l, dtype=np.float64).reshape(self.nu), np.asarray(h, dtype=np.float64).reshape(self.nu)
        if np.any(lo != lo[0, 0]) or np.any(hi != hi[0, 0]):
            check(lib().nmpc_b200_ddp_set_input_limits_mpc(self._h, n_ticks, N, lo.ctypes.data_as(C.c_void_p),
                                                           hi.ctypes.data_as(C.c_void_p)))

    def solve(self, current_t, current_x, initial_u_list):
        """Single-instance ``solve`` (DDPSolver.hpp:27-141); returns True iff converged (retval == 1)."""
        u = np.asarray(initial_u_list, dtype=np.float64)
        n_steps = u.shape[0] if u.ndim >= 1 else 0
        u = u.reshape(1, n_steps, -1) if u.size else u.reshape(1, n_steps, self.nu)
        ok = self.solve_batch(current_t, np.asarray(current_x, dtype=np.float64).reshape(1, self.nx), u)
        return bool(ok[0])

    def solve_batch(self, current_t, x0, u_init, stream=None, read_status=True):
        """Solve B independent instances.  x0: [B, NX]; u_init: [B, n_steps, NU] (numpy, or float64 torch
        CUDA tensors for a device-resident call).  Returns the per-instance ``solve()`` return values."""
        self._apply_config()
        self._apply_limits(float(current_t))
        x_shape = tuple(x0.shape)
        if len(x_shape) != 2 or x_shape[1] != self.nx:
            raise InvalidArgument(_capi.ERR_INVALID_ARGUMENT, f"x0 must be [B, {self.nx}], got {x_shape}")
        B = x_shape[0]
        u_shape = tuple(u_init.shape)
        if len(u_shape) != 3 or u_shape[0] != B or u_shape[2] != self.nu:
            raise InvalidArgument(_capi.ERR_INVALID_ARGUMENT, f"u_init must be [{B}, n_steps, {self.nu}], got {u_shape}")
        n_steps = u_shape[1]
        px, dev_x, keep_x = _capi.as_device_or_host(x0, (B, self.nx))
        pu, dev_u, keep_u = _capi.as_device_or_host(u_init, (B, n_steps, self.nu))
        if dev_x != dev_u:
            raise InvalidArgument(_capi.ERR_INVALID_ARGUMENT, "x0 and u_init must both be host or both be device arrays")
        check(lib().nmpc_b200_ddp_solve(self._h, B, float(current_t), px, pu, n_steps, int(dev_x),
                                        self._stream_ptr(stream)))
        self._B = B
        self._keep = (keep_x, keep_u)
        if not read_status:
            return None
        return self.status() == 1

    def run_mpc(self, current_t, x0, u_init, n_ticks, tick_dt, plant="model", shift_inputs=True, clamp_u0=False,
                sim_dt=None, n_substeps=1, stream=None):
        """The reference's receding-horizon loops for B instances, every tick on the device (no host round trip):
        solve -> apply u_list[0] -> advance current_x -> warm-start the next solve.

        ``plant="model"``: current_x <- x_list[1], ``shift_inputs=True`` (TestDDPBipedal.cpp:262-267);
        ``plant="sim"``: current_x <- stateEq(t, x, u, sim_dt) ``n_substeps`` times, usually with
        ``shift_inputs=False, clamp_u0=True`` (TestDDPCartPole.cpp:330, :388-396).
        Returns a dict: x [B, n_ticks+1, NX], u [B, n_ticks, NU] (applied inputs), iters, status [B, n_ticks]."""
        self._apply_config()
        self._apply_limits(float(current_t))
        self._apply_limits_mpc(float(current_t), int(n_ticks), float(tick_dt))
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        u_init = np.ascontiguousarray(u_init, dtype=np.float64)
        B = x0.shape[0]
        if x0.shape != (B, self.nx) or u_init.ndim != 3 or u_init.shape[0] != B or u_init.shape[2] != self.nu:
            raise InvalidArgument(_capi.ERR_INVALID_ARGUMENT, f"x0 must be [B, {self.nx}] and u_init [B, n_steps, {self.nu}]")
        mc = _capi.MpcConfigStruct()
        mc.n_ticks = int(n_ticks)
        mc.plant = {"model": 0, "sim": 1}[plant]
        mc.shift_inputs = int(bool(shift_inputs))
        mc.clamp_u0 = int(bool(clamp_u0))
        mc.n_substeps = int(n_substeps)
        mc.tick_dt = float(tick_dt)
        mc.sim_dt = float(tick_dt if sim_dt is None else sim_dt)
        T = max(int(n_ticks), 0)
        out = {"x": np.empty((B, T + 1, self.nx)), "u": np.empty((B, T, self.nu)),
               "iters": np.empty((B, T), dtype=np.int32), "status": np.empty((B, T), dtype=np.int32)}
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().nmpc_b200_ddp_run_mpc(self._h, B, float(current_t), vp(x0), vp(u_init), int(u_init.shape[1]),
                                          C.byref(mc), vp(out["x"]), vp(out["u"]), vp(out["iters"]), vp(out["status"]),
                                          0, self._stream_ptr(stream)))
        self._B = B
        return out

    def controlData(self):
        return ControlData(x_list=self._get_f64(F_X, (self._B, self._config.horizon_steps + 1, self.nx)),
                           u_list=self._get_f64(F_U, (self._B, self._config.horizon_steps, self.nu)),
                           cost_list=self._get_f64(F_COST_LIST, (self._B, self._config.horizon_steps + 1)))

    def traceDataList(self, instance=0):
        """traceDataList() of one instance (DDPSolver.h:294-297)."""
        tr = self.trace()[instance]
        n = int(self.n_trace()[instance])
        dur = self.iterationDurations()
        out = []
        for r in range(n):
            row = tr[r]
            d = dur[r] if r < len(dur) else (0.0, 0.0, 0.0, 0.0)
            out.append(TraceData(iter=int(row[0]), cost=row[1], lambda_=row[2], dlambda=row[3], alpha=row[4],
                                 k_rel_norm=row[5], cost_update_actual=row[6], cost_update_expected=row[7],
                                 cost_update_ratio=row[8], duration_derivative=float(d[0]),
                                 duration_backward=float(d[1]), duration_forward=float(d[2] + d[3])))
        return out

    def iterationDurations(self):
        """[rows][4] ms per trace entry: derivative, backward, first line-search candidate, other candidates (the
        stages of the whole batch; zeros unless enable_timing(True) was set before the solve)."""
        rows = self._config.max_iter + 1
        ms = np.zeros((rows, 4))
        n = C.c_int(0)
        check(lib().nmpc_b200_ddp_get_iteration_durations(self._h, ms.ctypes.data_as(C.c_void_p), rows, C.byref(n)))
        return ms[:n.value]

    def dumpTraceDataList(self, file_path, instance=0):
        """Same 12-column, space-separated table as DDPSolver::dumpTraceDataList (DDPSolver.hpp:563-598)."""
        with open(file_path, "w") as f:
            f.write("iter cost lambda dlambda alpha k_rel_norm cost_update_actual cost_update_expected "
                    "cost_update_ratio duration_derivative duration_backward duration_forward\n")
            for t in self.traceDataList(instance):
                f.write(f"{t.iter} {t.cost:g} {t.lambda_:g} {t.dlambda:g} {t.alpha:g} {t.k_rel_norm:g} "
                        f"{t.cost_update_actual:g} {t.cost_update_expected:g} {t.cost_update_ratio:g} "
                        f"{t.duration_derivative:g} {t.duration_backward:g} {t.duration_forward:g}\n")

    def computationDuration(self):
        """ComputationDuration (DDPSolver.h:219-247) in ms from CUDA events; needs enable_timing(True)."""
        ms = (C.c_double * 8)()
        launches = (C.c_int * 4)()
        check(lib().nmpc_b200_ddp_get_durations(self._h, ms, launches))
        return {
            "solve": ms[0], "setup": ms[1], "opt": ms[2], "derivative": ms[3], "backward": ms[4], "forward": ms[5],
            "copy_in": ms[6], "copy_out": ms[7],
            "launches": {"rollout": launches[0], "derivative": launches[1], "backward": launches[2],
                         "forward": launches[3]},
        }

    # -- batch accessors ------------------------------------------------------------------------
    def enable_timing(self, enable=True):
        self._timing = bool(enable)
        check(lib().nmpc_b200_ddp_enable_timing(self._h, int(self._timing)))

    def synchronize(self):
        check(lib().nmpc_b200_ddp_sync(self._h))

    def status(self):
        return self._get_i32(F_STATUS)

    def iterations(self):
        return self._get_i32(F_ITERS)

    def n_forward(self):
        return self._get_i32(F_N_FORWARD)

    def n_backward(self):
        return self._get_i32(F_N_BACKWARD)

    def n_trace(self):
        return self._get_i32(F_N_TRACE)

    def cost(self):
        return self._get_f64(F_COST, (self._B,))

    def u0(self, out=None, stream=None):
        """First-step controls u_list[0] of every instance, [B, NU] (what an MPC loop applies)."""
        return self._get_f64(F_U0, (self._B, self.nu), out=out, stream=stream)

    def k_list(self):
        return self._get_f64(F_K_FF, (self._B, self._config.horizon_steps, self.nu))

    def K_list(self):
        """[B, N, NU, NX]"""
        raw = self._get_f64(F_K_FB, (self._B, self._config.horizon_steps, self.nx, self.nu))
        return raw.transpose(0, 1, 3, 2).copy()

    def trace(self):
        """[B, max_iter+1, 9] with columns TRACE_FIELDS; rows past n_trace are zero."""
        return self._get_f64(F_TRACE, (self._B, self._config.max_iter + 1, len(TRACE_FIELDS)))

    def get_into(self, what, out, stream=None):
        """Copy field `what` into a preallocated numpy array or float64 torch CUDA tensor."""
        return self._get_f64(what, tuple(out.shape), out=out, stream=stream)

    def set_tuning(self, **knobs):
        """Pin kernel-selection knobs (nmpc_b200_ddp_set_tuning), e.g. ``set_tuning(backward_lanes=0, forward_split=0)``."""
        for key, value in knobs.items():
            check(lib().nmpc_b200_ddp_set_tuning(self._h, key.encode(), int(value)))

    def get_tuning(self, key):
        v = C.c_int()
        check(lib().nmpc_b200_ddp_get_tuning(self._h, key.encode(), C.byref(v)))
        return v.value

    def get_to_device_ptr(self, what, ptr, nbytes, stream=None):
        """Field `what` into device memory given as a raw address -- e.g. a row of a sharding.PeerBuffer that lives on
        another GPU / in another process: the gather kernel stores straight into it."""
        check(lib().nmpc_b200_ddp_get(self._h, int(what), C.c_void_p(int(ptr)), int(nbytes), 1, self._stream_ptr(stream)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().nmpc_b200_ddp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover
            pass

    # -- helpers ----------------------------------------------------------------------------------
    @staticmethod
    def _stream_ptr(stream):
        if stream is None:
            return None
        if hasattr(stream, "cuda_stream"):
            return C.c_void_p(stream.cuda_stream)
        return C.c_void_p(int(stream))

    def _apply_config(self):
        s = self._config.to_struct()
        raw = bytes(s)
        if raw != self._applied:
            check(lib().nmpc_b200_ddp_set_config(self._h, C.byref(s)))
            self._applied = raw

    def _get_f64(self, what, shape, out=None, stream=None):
        if self._B <= 0:
            raise _capi.NmpcB200Error(_capi.ERR_RUNTIME, "no solve() yet")
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        ptr, on_dev, keep = _capi.as_device_or_host(out, shape)
        if not on_dev and keep.ctypes.data != out.ctypes.data:
            raise InvalidArgument(_capi.ERR_INVALID_ARGUMENT, "out must be a contiguous float64 array")
        nbytes = int(np.prod(shape)) * 8
        check(lib().nmpc_b200_ddp_get(self._h, int(what), ptr, nbytes, int(on_dev), self._stream_ptr(stream)))
        return out

    def _get_i32(self, what):
        if self._B <= 0:
            raise _capi.NmpcB200Error(_capi.ERR_RUNTIME, "no solve() yet")
        out = np.empty(self._B, dtype=np.int32)
        check(lib().nmpc_b200_ddp_get(self._h, int(what), out.ctypes.data_as(C.c_void_p), out.nbytes, 0, None))
        return out
