#!/usr/bin/env python
"""Generate tests/golden/reference_centroidal.npz with the REFERENCE's own DDPSolver<9, Eigen::Dynamic>
(oracle/_ref/libnmpc_ref.so: /root/reference's DDPSolver.h(.hpp) compiled unmodified against oracle/ref/eigen_shim)
on the problem and MPC loop of nmpc_ddp/tests/src/TestDDPCentroidalMotion.cpp: n_x = 9, 16 ridge-force inputs while a
foot is in contact and NO input during the flight phase 1.4 s <= t < 1.6 s, horizon 3 s / 0.03 s = 100 steps,
max_iter 500 for the first solve and 3 afterwards, 100 MPC ticks (the whole test).

  x_first, u_first, cost_first   the first solve (full trajectories, u padded to 16 with zeros)
  x_log, u0_log, dim_log, iters_log   per tick: current_x, u_list[0] (padded), its size, iterations
  x, u                           trajectories of the last solve

Only runnable where /root/reference exists.   python tests/golden/make_golden_centroidal.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_lib as R  # noqa: E402

N, TICKS = 100, 100
cases = dict(R.centroidal_mpc(N, TICKS))
cases["N"] = np.array(N)
cases["ticks"] = np.array(TICKS)
np.savez_compressed(os.path.join(HERE, "reference_centroidal.npz"), **cases)
print("wrote reference_centroidal.npz:", {k: v.shape for k, v in cases.items()})
print("iters:", cases["iters_log"][:12], "dims:", sorted(set(cases["dim_log"].tolist())), "cost_first", cases["cost_first"])
print("final x:", cases["x_log"][-1])
