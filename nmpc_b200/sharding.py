"""Batch sharding across the GPUs of one box.

Instances never interact (the reference runs one DDPSolver object per problem,
nmpc_ddp/include/nmpc_ddp/DDPSolver.h:329-374), so the batch splits into contiguous chunks with no
data-path collective.  The only exchange is the optional gather of first-step controls u_list[0]
(what an MPC loop applies, e.g. nmpc_ddp/tests/src/TestDDPBipedal.cpp:254).  Three ways to run it, all over the C ABI
(include/nmpc_b200/c_api.h, "several GPUs, one box"):

  * ShardedDDPSolver -- ONE process, a solver handle + stream + host worker per device (nmpc_b200_ddp_create_sharded);
  * one process per GPU (torchrun) with PeerBuffer: every rank's gather kernel stores its u0 rows straight into rank 0's
    device buffer over NVLink and raises a flag word; no collective in the step (nmpc_b200_peer_*);
  * one process per GPU with gather_first_controls: the NCCL / gloo all-gather (kept for comparison and for CPU tests).
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import check, lib


def shard_range(total, world_size, rank):
    """Contiguous [begin, end) of `total` instances owned by `rank`: floor(total / world) each, the
    remainder spread one per rank over the lowest ranks."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, rem = divmod(int(total), int(world_size))
    begin = rank * base + min(rank, rem)
    end = begin + base + (1 if rank < rem else 0)
    return begin, end


def shard_sizes(total, world_size):
    return [shard_range(total, world_size, r)[1] - shard_range(total, world_size, r)[0] for r in range(world_size)]


def gather_first_controls(u0_local, total, group=None):
    """All-gather of per-rank first-step controls [B_rank, NU] into [total, NU] on every rank.

    Works for ragged shards (pads to the largest shard).  `u0_local` is a torch tensor on the device
    the process group communicates on (CUDA for NCCL, CPU for gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    nu = u0_local.shape[1]
    if len(set(sizes)) == 1:
        out = torch.empty((total, nu), dtype=u0_local.dtype, device=u0_local.device)
        dist.all_gather_into_tensor(out, u0_local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad, nu), dtype=u0_local.dtype, device=u0_local.device)
    buf[:u0_local.shape[0]] = u0_local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


class ShardedDDPSolver:
    """A batch of DDPSolver objects sharded over several GPUs from one process (nmpc_b200_ddp_create_sharded)."""

    def __init__(self, problem, params=None, total_capacity=1, devices=None, config=None):
        from .ddp import DDPConfiguration

        self._h = C.c_void_p()
        self.problem = problem
        self.nx, self.nu, self.ng, self.n_params = _capi.model_dims(problem)
        self.params = (_capi.model_default_params(problem) if params is None else np.ascontiguousarray(
            params, dtype=np.float64))
        self._config = config if config is not None else DDPConfiguration()
        dev = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        st = self._config.to_struct()
        check(lib().nmpc_b200_ddp_create_sharded(problem.encode(), self.params.ctypes.data_as(C.c_void_p),
                                                 int(self.params.size), C.byref(st), int(total_capacity),
                                                 None if dev is None else dev.ctypes.data_as(C.c_void_p),
                                                 0 if dev is None else int(dev.size), C.byref(self._h)))
        self._applied = bytes(st)
        self._B = 0

    def config(self):
        return self._config

    def num_shards(self):
        return lib().nmpc_b200_ddp_sharded_num_shards(self._h)

    def shard_range(self, B, shard):
        """(begin, end, device) of `shard` for a solve with B instances."""
        b, e, d = C.c_int(), C.c_int(), C.c_int()
        check(lib().nmpc_b200_ddp_sharded_range(self._h, int(B), int(shard), C.byref(b), C.byref(e), C.byref(d)))
        return b.value, e.value, d.value

    def _apply_config(self):
        st = self._config.to_struct()
        if bytes(st) != self._applied:
            check(lib().nmpc_b200_ddp_sharded_set_config(self._h, C.byref(st)))
            self._applied = bytes(st)

    def setInputLimitsFunc(self, input_limits):
        """Constant (lower, upper) limits for every shard (DDPSolver.h:282-285)."""
        self._apply_config()
        lo = np.ascontiguousarray(input_limits[0], dtype=np.float64).reshape(self.nu)
        hi = np.ascontiguousarray(input_limits[1], dtype=np.float64).reshape(self.nu)
        check(lib().nmpc_b200_ddp_sharded_set_input_limits(self._h, lo.ctypes.data_as(C.c_void_p),
                                                           hi.ctypes.data_as(C.c_void_p)))

    def solve_batch(self, current_t, x0, u_init):
        """DDPSolver::solve for every instance: x0 [B, NX], u_init [B, N, NU] (host arrays).  Returns solve()'s bool per
        instance."""
        self._apply_config()
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, self.nx)
        B = x0.shape[0]
        u_init = np.ascontiguousarray(u_init, dtype=np.float64)
        if u_init.ndim != 3 or u_init.shape[0] != B or u_init.shape[2] != self.nu:
            raise ValueError(f"u_init must be [B={B}, N, NU={self.nu}]")
        check(lib().nmpc_b200_ddp_sharded_solve(self._h, B, float(current_t), x0.ctypes.data_as(C.c_void_p),
                                                u_init.ctypes.data_as(C.c_void_p), int(u_init.shape[1])))
        self._B = B
        return self.get(6, (B,), np.int32) == 1

    def get(self, what, shape, dtype=np.float64, out=None, dst_device=-1):
        """Field `what` (nmpc_b200_ddp_field) of the last solve over all shards.  `out`: a host numpy array, or -- with
        dst_device >= 0 -- a torch CUDA tensor on that device, which every shard stores into directly."""
        if out is None:
            out = np.zeros(shape, dtype=dtype)
        if isinstance(out, np.ndarray):
            ptr, nbytes = out.ctypes.data_as(C.c_void_p), out.nbytes
        else:
            ptr, nbytes = C.c_void_p(out.data_ptr()), out.numel() * out.element_size()
        check(lib().nmpc_b200_ddp_sharded_get(self._h, int(what), ptr, nbytes, int(dst_device)))
        return out

    def u_list(self):
        return self.get(1, (self._B, self._config.horizon_steps, self.nu))

    def u0(self, out=None, dst_device=-1):
        return self.get(11, (self._B, self.nu), out=out, dst_device=dst_device)

    def cost(self):
        return self.get(10, (self._B,))

    def iterations(self):
        return self.get(7, (self._B,), np.int32)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().nmpc_b200_ddp_sharded_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover
            pass


class PeerBuffer:
    """A device buffer of rank `owner` that every process of the box stores into (nmpc_b200_peer_*): `n_rows` rows of
    `row_bytes` followed by one flag word per rank and a time-out word.

    Create on every rank of an initialised torch.distributed group (any backend; the 64-byte handle travels through
    broadcast_object_list once, outside any timed region)."""

    def __init__(self, n_rows, row_bytes, device, owner=0, group=None):
        import torch.distributed as dist

        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device, self.owner = int(device), int(owner)
        self.data_bytes = (int(n_rows) * int(row_bytes) + 255) // 256 * 256
        self.row_bytes = int(row_bytes)
        total = self.data_bytes + 8 * (self.world + 1)
        self._ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        if self.rank == owner:
            check(lib().nmpc_b200_peer_buffer_create(total, self.device, C.byref(self._ptr), handle))
        box = [bytes(handle) if self.rank == owner else None]
        dist.broadcast_object_list(box, src=owner, group=group)
        if self.rank != owner:
            handle = (C.c_ubyte * 64).from_buffer_copy(box[0])
            check(lib().nmpc_b200_peer_buffer_open(handle, self.device, C.byref(self._ptr)))
        self._step = 0

    @property
    def ptr(self):
        return self._ptr.value

    def row_ptr(self, row):
        return self._ptr.value + int(row) * self.row_bytes

    @property
    def flags_ptr(self):
        return self._ptr.value + self.data_bytes

    def signal(self, value, stream=None):
        """After this rank's stores on `stream`: flag[rank] <- value."""
        check(lib().nmpc_b200_peer_signal(C.c_void_p(self.flags_ptr + 8 * self.rank), int(value), self.device,
                                          C.c_void_p(stream or 0)))

    def wait(self, value, stream=None, timeout_ms=2000):
        """Owner only: `stream` continues when every rank's flag has reached `value`."""
        check(lib().nmpc_b200_peer_wait(C.c_void_p(self.flags_ptr), self.world, int(value), int(timeout_ms),
                                        self.device, C.c_void_p(stream or 0)))

    def check(self, stream=None):
        check(lib().nmpc_b200_peer_check(C.c_void_p(self.flags_ptr), self.world, self.device, C.c_void_p(stream or 0)))

    def read(self, n_rows, dtype=np.float64):
        """Owner only: the first n_rows rows as a host array (test helper; synchronises the device)."""
        import torch

        n = int(n_rows) * self.row_bytes
        out = np.zeros(n // np.dtype(dtype).itemsize, dtype=dtype)
        torch.cuda.synchronize(self.device)
        from cuda.bindings import runtime as rt  # cuda-python

        (err,) = rt.cudaMemcpy(out.ctypes.data, self._ptr.value, n, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        if int(err) != 0:
            raise RuntimeError(f"cudaMemcpy failed: {err}")
        return out

    def close(self):
        if self._ptr:
            if self.rank == self.owner:
                lib().nmpc_b200_peer_buffer_destroy(self._ptr, self.device)
            else:
                lib().nmpc_b200_peer_buffer_close(self._ptr, self.device)
            self._ptr = C.c_void_p()


class ShardedFmpcSolver:
    """A batch of FmpcSolver objects sharded over several GPUs from one process (nmpc_b200_fmpc_create_sharded)."""

    def __init__(self, problem, params=None, total_capacity=1, devices=None, config=None):
        from .fmpc import FmpcConfiguration

        self._h = C.c_void_p()
        self.nx, self.nu, self.ng, self.n_params = _capi.model_dims(problem)
        self.params = (_capi.model_default_params(problem) if params is None else np.ascontiguousarray(
            params, dtype=np.float64))
        self._config = config if config is not None else FmpcConfiguration()
        dev = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        st = self._config.to_struct()
        check(lib().nmpc_b200_fmpc_create_sharded(problem.encode(), self.params.ctypes.data_as(C.c_void_p),
                                                  int(self.params.size), C.byref(st), int(total_capacity),
                                                  None if dev is None else dev.ctypes.data_as(C.c_void_p),
                                                  0 if dev is None else int(dev.size), C.byref(self._h)))
        self._applied = bytes(st)
        self._B = 0

    def config(self):
        return self._config

    def num_shards(self):
        return lib().nmpc_b200_fmpc_sharded_num_shards(self._h)

    def solve_batch(self, current_t, x0, var):
        """FmpcSolver::solve for every instance: x0 [B, NX] and a fmpc.Variable of host arrays.  Returns the status words."""
        st = self._config.to_struct()
        if bytes(st) != self._applied:
            check(lib().nmpc_b200_fmpc_sharded_set_config(self._h, C.byref(st)))
            self._applied = bytes(st)
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, self.nx)
        B = x0.shape[0]
        arrs = [np.ascontiguousarray(getattr(var, n), dtype=np.float64)
                for n in ("x_list", "u_list", "lambda_list", "s_list", "nu_list")]
        n_steps = int(arrs[1].shape[1])
        check(lib().nmpc_b200_fmpc_sharded_solve(self._h, B, float(current_t), x0.ctypes.data_as(C.c_void_p),
                                                 *[a.ctypes.data_as(C.c_void_p) for a in arrs], n_steps))
        self._B = B
        return self.get(8, (B,), np.int32)

    def get(self, what, shape, dtype=np.float64):
        out = np.zeros(shape, dtype=dtype)
        check(lib().nmpc_b200_fmpc_sharded_get(self._h, int(what), out.ctypes.data_as(C.c_void_p), out.nbytes, -1))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().nmpc_b200_fmpc_sharded_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pragma: no cover
            pass
