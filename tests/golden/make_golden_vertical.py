#!/usr/bin/env python
"""Generate tests/golden/reference_vertical.npz with the REFERENCE's own DDPSolver<2, Eigen::Dynamic>
(oracle/_ref/libnmpc_ref.so: /root/reference's DDPSolver.h(.hpp), BoxQP.h compiled unmodified against
oracle/ref/eigen_shim) on the problem and MPC loop of nmpc_ddp/tests/src/TestDDPVerticalMotion.cpp: time-varying
input dimension (1, 2, 1, 0, 1 contact forces along the time axis), horizon 3 s / 0.01 s, initial_lambda 1e-6,
max_iter 500 for the first solve and 3 afterwards, with and without input limits [0, 30] N.

  first/*   the first solve only (full trajectories)
  loop/*    230 ticks: the terminal time crosses 4.5 s (tick 150, dimension 1 -> 0) and 5.0 s (tick 200, 0 -> 1),
            so both branches of the test's warm-start rule (:306-315) are exercised

Only runnable where /root/reference exists.   python tests/golden/make_golden_vertical.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_lib as R  # noqa: E402

N, TICKS = 300, 230
cases = {}
for wc in (0, 1):
    first = R.vertical_mpc(N, wc, 1)
    loop = R.vertical_mpc(N, wc, TICKS)
    tag = "box" if wc else "free"
    for k in ("x", "u", "iters_log", "dim_log"):
        cases[f"first_{tag}/{k}"] = first[k]
    for k in ("x_log", "u0_log", "dim_log", "iters_log", "x", "u"):
        cases[f"loop_{tag}/{k}"] = loop[k]
cases["N"] = np.array(N)
cases["ticks"] = np.array(TICKS)
np.savez_compressed(os.path.join(HERE, "reference_vertical.npz"), **cases)
print("wrote reference_vertical.npz:", {k: v.shape for k, v in cases.items()})
