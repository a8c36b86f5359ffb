#!/usr/bin/env python
"""Wall-clock of BASELINE.json configs[2] (FMPC cart-pole, B=1024) and configs[3] (quadrotor iLQR fp32, B=8192)
on one GPU: host buffers in, first controls out, best of 5 after 2 warm-ups.  Parity of the same runs is the job
of tests/; this tool only reports time (gpurun_out/time_configs.json).

    python tools/time_configs.py [quadrotor|fmpc|fmpc_sweep|centroidal|centroidal_profile|all]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nmpc_b200  # noqa: E402
import oracle_lib as O  # noqa: E402  (synthetic inputs only)


def best_of(fn, sync, n=5, warm=2):
    for _ in range(warm):
        fn()
        sync()
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        fn()
        sync()
        ts.append(time.perf_counter() - t)
    return min(ts)


def quadrotor():
    from test_quadrotor_gpu import N, hover_inputs, quadrotor_x0

    B = 8192
    p = O.default_params("quadrotor")
    x0, u0 = quadrotor_x0(B, 4), hover_inputs(B)
    out = {}
    for name in ("quadrotor", "quadrotor_f64"):
        s = nmpc_b200.DDPSolver(name, params=p, batch_capacity=B)
        c = s.config()
        c.horizon_steps, c.max_iter, c.k_rel_norm_thre, c.cost_update_thre = N, 10, 0.0, 0.0
        t = best_of(lambda: s.solve_batch(0.0, x0, u0, read_status=False), s.synchronize)
        s.enable_timing(True)
        s.solve_batch(0.0, x0, u0, read_status=False)
        d = s.computationDuration()
        out[name] = {"batch": B, "horizon": N, "iters": 10, "ms": 1e3 * t, "traj_per_s": B / t,
                     "stage_ms": {k: d[k] for k in ("derivative", "backward", "forward", "setup", "solve")},
                     "fwd_passes_mean": float(s.n_forward().mean()), "bwd_passes_mean": float(s.n_backward().mean()),
                     "bwd_passes_max": int(s.n_backward().max()), "status_counts": {int(k): int(v) for k, v in zip(
                         *np.unique(s.status(), return_counts=True))}}
        s.close()
    return out


def centroidal():
    """First solve of TestDDPCentroidalMotion (n_x = 9, n_u = 16 / 0, N = 100, reference termination rules) for a batch
    of perturbed starts.  Not a BASELINE.json config: reported so that the cost of the untuned 9 x 16 kernels is on
    record."""
    N = 100
    p = O.default_params("centroidal_motion")
    out = {}
    for B in (64, 1024):
        x0 = np.zeros((B, 9))
        x0[:, 2] = 1.0
        x0[:, :3] += np.random.default_rng(B).uniform(-0.05, 0.05, (B, 3))
        u0 = np.zeros((B, N, 16))
        s = nmpc_b200.DDPSolver("centroidal_motion", params=p, batch_capacity=B)
        s.config().horizon_steps = N
        t = best_of(lambda: s.solve_batch(0.0, x0, u0, read_status=False), s.synchronize, n=3, warm=1)
        s.enable_timing(True)
        s.solve_batch(0.0, x0, u0, read_status=False)
        d = s.computationDuration()
        out[f"centroidal_B{B}"] = {"batch": B, "horizon": N, "ms": 1e3 * t, "traj_per_s": B / t,
                                   "iters_mean": float(s.iterations().mean()), "iters_max": int(s.iterations().max()),
                                   "stage_ms": {k: d[k] for k in ("derivative", "backward", "forward", "setup", "solve")},
                                   "status_counts": {int(k): int(v) for k, v in zip(*np.unique(s.status(), return_counts=True))}}
        s.close()
    return out


def centroidal_profile(B=1024):
    """Four full iterations (termination thresholds off) of the centroidal problem, three times: every K2 launch is
    live, for `ncu -k regex:backward_wide_kernel --launch-skip 5 --launch-count 1` (tools/profile_kernels.sh)."""
    N = 100
    p = O.default_params("centroidal_motion")
    x0 = np.zeros((B, 9))
    x0[:, 2] = 1.0
    x0[:, :3] += np.random.default_rng(B).uniform(-0.05, 0.05, (B, 3))
    u0 = np.zeros((B, N, 16))
    s = nmpc_b200.DDPSolver("centroidal_motion", params=p, batch_capacity=B)
    c = s.config()
    c.horizon_steps, c.max_iter, c.k_rel_norm_thre, c.cost_update_thre = N, 4, 0.0, 0.0
    t = best_of(lambda: s.solve_batch(0.0, x0, u0, read_status=False), s.synchronize, n=2, warm=1)
    return {"centroidal_profile": {"batch": B, "iters": 4, "ms": 1e3 * t, "bwd_passes_mean": float(s.n_backward().mean())}}


def fmpc(B=1024):
    N = 100
    x0 = O.cartpole_x0(B, 3)
    s = nmpc_b200.FmpcSolver("cartpole", batch_capacity=B)
    s.config().horizon_steps = N
    s.config().max_iter = 10
    var = s.make_variable(B)
    var.reset(0.0, 0.0, 0.0, 1.0, 1.0)
    t = best_of(lambda: s.solve_batch(0.0, x0, var), s.synchronize)
    s.enable_timing(True)
    s.solve_batch(0.0, x0, var)
    d = s.computationDuration()
    return {"fmpc_cartpole": {"batch": B, "horizon": N, "iters": 10, "ms": 1e3 * t, "solves_per_s": B / t,
                              "stage_ms": {k: v for k, v in d.items() if isinstance(v, float)}}}


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    res = {"env": {k: v for k, v in os.environ.items() if k.startswith("NMPC_B200_")}}
    if what in ("quadrotor", "all"):
        res.update(quadrotor())
    if what in ("fmpc", "all"):
        res.update(fmpc())
    if what == "centroidal":
        res.update(centroidal())
    if what == "centroidal_profile":
        res.update(centroidal_profile())
    if what == "fmpc_sweep":
        res["fmpc_sweep"] = [fmpc(B)["fmpc_cartpole"] for B in (1024, 4096, 16384, 65536)]
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "time_configs.json"), "a") as f:
        f.write(json.dumps(res) + "\n")
