#!/usr/bin/env python
"""Closed-loop throughput of the device-resident MPC loop (nmpc_b200_ddp_run_mpc): instance-ticks per second for a batch
of warm-started cart-pole MPCs (TestDDPCartPole's settings: horizon 2 s / 0.01 s, max_iter 3, limits +-15 N, tick 4 ms,
plant at 2 ms), 100 ticks after 20 warm-up ticks.  Appends to gpurun_out/time_mpc.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nmpc_b200  # noqa: E402
import oracle_lib as O  # noqa: E402  (synthetic initial states only)

rows = []
for B in [int(a) for a in sys.argv[1:]] or [16, 1024, 4096, 32768]:
    N = 200
    x0 = O.cartpole_x0(B, B)
    s = nmpc_b200.DDPSolver("cartpole", batch_capacity=B)
    c = s.config()
    c.horizon_steps, c.max_iter, c.with_input_constraint = N, 3, True
    s.setInputLimitsFunc((np.array([-15.0]), np.array([15.0])))
    kw = dict(tick_dt=0.004, plant="sim", shift_inputs=False, clamp_u0=True, sim_dt=0.002, n_substeps=2)
    warm = s.run_mpc(0.0, x0, np.zeros((B, N, 1)), n_ticks=20, **kw)
    u = s.controlData().u_list
    t = time.perf_counter()
    log = s.run_mpc(20 * 0.004, warm["x"][:, -1], u, n_ticks=100, **kw)
    el = time.perf_counter() - t
    rows.append({"batch": B, "horizon": N, "max_iter": 3, "ticks": 100, "ms_per_tick": 1e3 * el / 100,
                 "instance_ticks_per_s": B * 100 / el, "mean_iterations": float(log["iters"].mean())})
    print(json.dumps(rows[-1]), flush=True)
    s.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "time_mpc.json"), "w") as f:
    json.dump(rows, f, indent=1)
