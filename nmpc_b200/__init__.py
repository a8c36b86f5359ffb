"""nmpc_b200 -- batched DDP/iLQR and FMPC solves on NVIDIA B200 (sm_100a).

Python host-side mirror of the reference's solver API (isri-aist/NMPC ``nmpc_ddp::DDPSolver``,
``nmpc_fmpc::FmpcSolver``) on top of the C ABI of ``libnmpc_b200.so``.  There is no CPU fallback:
importing works without a GPU (so the library's symbols can be checked), creating a solver does not.
"""
from ._capi import (NmpcB200Error, device_count, lib, library_path, load_plugin, model_default_params, model_dims,
                    model_eval, model_names)
from .ddp import DDPConfiguration, DDPSolver, ControlData, TraceData
from .fmpc import FmpcConfiguration, FmpcSolver, FmpcStatus, Variable

__all__ = [
    "NmpcB200Error", "device_count", "lib", "library_path", "load_plugin", "model_default_params", "model_dims", "model_eval",
    "model_names", "DDPConfiguration", "DDPSolver", "ControlData", "TraceData", "FmpcConfiguration", "FmpcSolver",
    "FmpcStatus", "Variable",
]
