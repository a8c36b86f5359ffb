"""Python mirror of ``nmpc_fmpc::FmpcSolver`` for batches of independent instances on one GPU.

Names follow the reference (isri-aist/NMPC nmpc_fmpc/include/nmpc_fmpc/FmpcSolver.h): ``Variable``
(x_list, u_list, lambda_list, s_list, nu_list; ``reset``), ``Configuration``, ``Status``,
``solve(current_t, current_x, initial_variable)``, ``variable()``, ``coeffList()`` gains,
``traceDataList()``.  All compute happens in libnmpc_b200.so.
"""
import ctypes as C
import enum
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import FmpcConfigStruct, InvalidArgument, check, lib

F_X, F_U, F_LAMBDA, F_S, F_NU, F_K_FF, F_K_FB, F_TRACE, F_STATUS, F_N_TRACE, F_U0 = range(11)

TRACE_FIELDS = ("iter", "kkt_error", "barrier_eps", "alpha_s", "alpha_nu")


class FmpcStatus(enum.IntEnum):
    """FmpcSolver::Status (FmpcSolver.h:92-114)."""
    Uninitialized = 0
    Succeeded = 1
    ErrorInForward = 2
    ErrorInBackward = 3
    ErrorInUpdate = 4
    MaxIterationReached = 5
    IterationContinued = 6


@dataclass
class FmpcConfiguration:
    """FmpcSolver::Configuration (FmpcSolver.h:58-89), same names and defaults."""
    print_level: int = 1
    horizon_steps: int = 100
    max_iter: int = 10
    kkt_error_thre: float = 1e-4
    check_nan: bool = True
    init_complementary_variable: bool = False
    update_barrier_eps: bool = True
    break_if_llt_fails: bool = False
    enable_line_search: bool = False
    merit_const_scale_from_lagrange_multipliers: bool = False
    initial_barrier_eps: float = 1e-4  # value of the barrier_eps_ member on entry (FmpcSolver.h:413-414)

    def to_struct(self):
        s = FmpcConfigStruct()
        s.horizon_steps = int(self.horizon_steps)
        s.max_iter = int(self.max_iter)
        for name in ("check_nan", "init_complementary_variable", "update_barrier_eps", "break_if_llt_fails",
                     "enable_line_search", "merit_const_scale_from_lagrange_multipliers"):
            setattr(s, name, int(bool(getattr(self, name))))
        s.kkt_error_thre = float(self.kkt_error_thre)
        s.initial_barrier_eps = float(self.initial_barrier_eps)
        return s


class Variable:
    """FmpcSolver::Variable (FmpcSolver.h:117-158) for a batch: arrays with a leading batch dimension."""

    def __init__(self, horizon_steps=0, batch=1, nx=0, nu=0, ng=0):
        self.horizon_steps = horizon_steps
        self.x_list = np.zeros((batch, horizon_steps + 1, nx))
        self.u_list = np.zeros((batch, horizon_steps, nu))
        self.lambda_list = np.zeros((batch, horizon_steps + 1, nx))
        self.s_list = np.zeros((batch, horizon_steps, ng))
        self.nu_list = np.zeros((batch, horizon_steps, ng))

    def reset(self, _x, _u, _lambda, _s, _nu):
        """Variable::reset (FmpcSolver.hpp:41-68)."""
        self.x_list[...] = _x
        self.u_list[...] = _u
        self.lambda_list[...] = _lambda
        self.s_list[...] = _s
        self.nu_list[...] = _nu

    def containsNaN(self):
        return any(not np.isfinite(a).all() for a in (self.x_list, self.u_list, self.lambda_list, self.s_list,
                                                      self.nu_list))


class FmpcSolver:
    """A batch of ``nmpc_fmpc::FmpcSolver`` objects sharing one problem functor, resident on one B200."""

    def __init__(self, problem, params=None, batch_capacity=1, device=0, config=None):
        self._h = C.c_void_p()
        self.problem = problem
        self.nx, self.nu, self.ng, self.n_params = _capi.model_dims(problem)
        self.params = (_capi.model_default_params(problem) if params is None else np.ascontiguousarray(
            params, dtype=np.float64))
        self.batch_capacity = int(batch_capacity)
        self._config = config if config is not None else FmpcConfiguration()
        self._B = 0
        s = self._config.to_struct()
        check(lib().nmpc_b200_fmpc_create(problem.encode(), self.params.ctypes.data_as(C.c_void_p),
                                          int(self.params.size), C.byref(s), self.batch_capacity, int(device),
                                          C.byref(self._h)))
        self._applied = bytes(s)

    def config(self):
        return self._config

    def make_variable(self, batch=1):
        return Variable(self._config.horizon_steps, batch, self.nx, self.nu, self.ng)

    def solve(self, current_t, current_x, initial_variable):
        """Single-instance solve (FmpcSolver.hpp:158-257): returns the FmpcStatus."""
        st = self.solve_batch(current_t, np.asarray(current_x, dtype=np.float64).reshape(1, self.nx),
                              initial_variable)
        return FmpcStatus(int(st[0]))

    def solve_batch(self, current_t, x0, var, stream=None):
        s = self._config.to_struct()
        raw = bytes(s)
        if raw != self._applied:
            check(lib().nmpc_b200_fmpc_set_config(self._h, C.byref(s)))
            self._applied = raw
        B = x0.shape[0]
        n_steps = var.u_list.shape[1]
        # checkVariable (FmpcSolver.hpp:288-312): sequence lengths
        N = self._config.horizon_steps
        for name, arr, want in (("x_list", var.x_list, N + 1), ("u_list", var.u_list, N),
                                ("lambda_list", var.lambda_list, N + 1), ("s_list", var.s_list, N),
                                ("nu_list", var.nu_list, N)):
            if arr.shape[1] != want:
                raise InvalidArgument(_capi.ERR_INVALID_ARGUMENT,
                                      f"[FMPC] {name} length should be {want} but {arr.shape[1]}.")
        keep = []
        ptrs = []
        devs = set()
        for arr, shape in ((x0, (B, self.nx)), (var.x_list, (B, N + 1, self.nx)), (var.u_list, (B, N, self.nu)),
                           (var.lambda_list, (B, N + 1, self.nx)), (var.s_list, (B, N, self.ng)),
                           (var.nu_list, (B, N, self.ng))):
            p, d, k = _capi.as_device_or_host(arr, shape)
            ptrs.append(p)
            keep.append(k)
            devs.add(d)
        if len(devs) != 1:
            raise InvalidArgument(_capi.ERR_INVALID_ARGUMENT, "all arrays must be host or all device")
        on_device = devs.pop()
        sp = None if stream is None else C.c_void_p(getattr(stream, "cuda_stream", stream))
        check(lib().nmpc_b200_fmpc_solve(self._h, B, float(current_t), *ptrs, n_steps, int(on_device), sp))
        self._B = B
        self._keep = keep
        return self.status()

    def run_mpc(self, current_t, x0, var, n_ticks, tick_dt, plant="sim", sim_dt=None, n_substeps=1, feedback=False,
                stream=None):
        """The reference's FMPC loops for B instances with every tick on the device: solve -> u_list[0] -> plant ->
        the Variable is the next warm start (TestFmpcOscillator.cpp:166-190); ``feedback=True`` adds
        K_0 (x_list[0] - current_x) at every plant sub-step (TestFmpcCartPole.cpp:351-356).  barrier_eps_ persists
        from tick to tick.  Returns dict x [B, T+1, NX], u [B, T, NU], kkt_error [B, T], status [B, T]."""
        s = self._config.to_struct()
        raw = bytes(s)
        if raw != self._applied:
            check(lib().nmpc_b200_fmpc_set_config(self._h, C.byref(s)))
            self._applied = raw
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        B = x0.shape[0]
        N = self._config.horizon_steps
        arrs = [x0] + [np.ascontiguousarray(a, dtype=np.float64) for a in (var.x_list, var.u_list, var.lambda_list,
                                                                           var.s_list, var.nu_list)]
        for name, arr, want in (("x_list", arrs[1], N + 1), ("u_list", arrs[2], N), ("lambda_list", arrs[3], N + 1),
                                ("s_list", arrs[4], N), ("nu_list", arrs[5], N)):
            if arr.shape[0] != B or arr.shape[1] != want:
                raise InvalidArgument(_capi.ERR_INVALID_ARGUMENT,
                                      f"[FMPC] {name} length should be {want} but {arr.shape[1]}.")
        mc = _capi.MpcConfigStruct()
        mc.n_ticks = int(n_ticks)
        mc.plant = {"model": 0, "sim": 1}[plant]
        mc.n_substeps = int(n_substeps)
        mc.feedback = int(bool(feedback))
        mc.tick_dt = float(tick_dt)
        mc.sim_dt = float(tick_dt if sim_dt is None else sim_dt)
        T = max(int(n_ticks), 0)
        out = {"x": np.empty((B, T + 1, self.nx)), "u": np.empty((B, T, self.nu)), "kkt_error": np.empty((B, T)),
               "status": np.empty((B, T), dtype=np.int32)}
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        sp = None if stream is None else C.c_void_p(getattr(stream, "cuda_stream", stream))
        check(lib().nmpc_b200_fmpc_run_mpc(self._h, B, float(current_t), *[vp(a) for a in arrs], int(arrs[2].shape[1]),
                                           C.byref(mc), vp(out["x"]), vp(out["u"]), vp(out["kkt_error"]),
                                           vp(out["status"]), 0, sp))
        self._B = B
        return out

    def variable(self):
        N = self._config.horizon_steps
        v = Variable(N, self._B, self.nx, self.nu, self.ng)
        v.x_list = self._get_f64(F_X, (self._B, N + 1, self.nx))
        v.u_list = self._get_f64(F_U, (self._B, N, self.nu))
        v.lambda_list = self._get_f64(F_LAMBDA, (self._B, N + 1, self.nx))
        v.s_list = self._get_f64(F_S, (self._B, N, self.ng))
        v.nu_list = self._get_f64(F_NU, (self._B, N, self.ng))
        return v

    def status(self):
        out = np.empty(self._B, dtype=np.int32)
        check(lib().nmpc_b200_fmpc_get(self._h, F_STATUS, out.ctypes.data_as(C.c_void_p), out.nbytes, 0, None))
        return out

    def n_trace(self):
        out = np.empty(self._B, dtype=np.int32)
        check(lib().nmpc_b200_fmpc_get(self._h, F_N_TRACE, out.ctypes.data_as(C.c_void_p), out.nbytes, 0, None))
        return out

    def trace(self):
        return self._get_f64(F_TRACE, (self._B, self._config.max_iter, len(TRACE_FIELDS)))

    def traceDataList(self, instance=0):
        """traceDataList() of one instance (FmpcSolver.h:232-251, :337-340) as a list of dicts."""
        tr = self.trace()[instance]
        n = int(self.n_trace()[instance])
        return [{"iter": int(tr[r, 0]), "kkt_error": float(tr[r, 1]), "barrier_eps": float(tr[r, 2]),
                 "alpha_s": float(tr[r, 3]), "alpha_nu": float(tr[r, 4]), "duration_coeff": 0.0,
                 "duration_backward": 0.0, "duration_forward": 0.0, "duration_update": 0.0} for r in range(n)]

    def dumpTraceDataList(self, file_path, instance=0):
        """Same 6-column, space-separated table as FmpcSolver::dumpTraceDataList (FmpcSolver.hpp:260-283), the format
        nmpc_fmpc/scripts/plotFmpcTraceData.py reads.  Per-iteration durations of a batched solve are not attributed
        to single instances: the duration columns are 0 (computationDuration() has the stage totals)."""
        with open(file_path, "w") as f:
            f.write("iter kkt_error duration_coeff duration_backward duration_forward duration_update\n")
            for t in self.traceDataList(instance):
                f.write(f"{t['iter']} {t['kkt_error']:g} {t['duration_coeff']:g} {t['duration_backward']:g} "
                        f"{t['duration_forward']:g} {t['duration_update']:g}\n")

    def k_list(self):
        return self._get_f64(F_K_FF, (self._B, self._config.horizon_steps, self.nu))

    def K_list(self):
        raw = self._get_f64(F_K_FB, (self._B, self._config.horizon_steps, self.nx, self.nu))
        return raw.transpose(0, 1, 3, 2).copy()

    def u0(self):
        return self._get_f64(F_U0, (self._B, self.nu))

    def enable_timing(self, enable=True):
        check(lib().nmpc_b200_fmpc_enable_timing(self._h, int(bool(enable))))

    def computationDuration(self):
        ms = (C.c_double * 8)()
        launches = (C.c_int * 4)()
        check(lib().nmpc_b200_fmpc_get_durations(self._h, ms, launches))
        return {"solve": ms[0], "setup": ms[1], "opt": ms[2], "coeff": ms[3], "backward": ms[4], "forward": ms[5],
                "update": ms[6], "copy": ms[7],
                "launches": {"coeff": launches[0], "backward": launches[1], "forward": launches[2],
                             "update": launches[3]}}

    def synchronize(self):
        check(lib().nmpc_b200_fmpc_sync(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().nmpc_b200_fmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover
            pass

    def _get_f64(self, what, shape):
        if self._B <= 0:
            raise _capi.NmpcB200Error(_capi.ERR_RUNTIME, "no solve() yet")
        out = np.empty(shape, dtype=np.float64)
        check(lib().nmpc_b200_fmpc_get(self._h, int(what), out.ctypes.data_as(C.c_void_p), out.nbytes, 0, None))
        return out
