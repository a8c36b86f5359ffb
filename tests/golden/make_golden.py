#!/usr/bin/env python
"""Generate tests/golden/reference_outputs.npz with the REFERENCE's own solver code.

Runs oracle/_ref/libnmpc_ref.so = /root/reference's nmpc_ddp/DDPSolver.h(.hpp), BoxQP.h and
nmpc_fmpc/FmpcSolver.h(.hpp), compiled unmodified against oracle/ref/eigen_shim (the image has no Eigen;
the shim restates Eigen's dense kernels), on the cases below.  Only runnable where /root/reference exists.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
import ref_lib as R  # noqa: E402

CASES = {}


def ddp_case(name, x0s, N, lo=None, hi=None, **cfg_kw):
    p = O.default_params("cartpole")
    cfg = R.ddp_config(horizon_steps=N, **cfg_kw)
    outs = [R.ddp_solve_cartpole(p, cfg, x0, np.zeros(N), u_lo=lo, u_hi=hi) for x0 in x0s]
    CASES[f"{name}/x0"] = np.array(x0s)
    CASES[f"{name}/N"] = np.array(N)
    CASES[f"{name}/cfg"] = np.array([cfg_kw.get("max_iter", 500), cfg_kw.get("with_input_constraint", 0),
                                     cfg_kw.get("k_rel_norm_thre", 1e-4), cfg_kw.get("cost_update_thre", 1e-7),
                                     cfg_kw.get("reg_type", 1)], dtype=np.float64)
    if lo is not None:
        CASES[f"{name}/limits"] = np.array([lo[0], hi[0]])
    for key in ("x", "u", "cost_list", "trace"):
        CASES[f"{name}/{key}"] = np.stack([o[key] for o in outs])
    CASES[f"{name}/n_trace"] = np.array([o["n_trace"] for o in outs])
    CASES[f"{name}/solve_ret"] = np.array([o["solve_ret"] for o in outs])


def fmpc_case(name, model, x0s, N, max_iter):
    p = O.default_params(model)
    nx, nu, ng, _ = O.model_dims(model)
    cfg = O.fmpc_config(horizon_steps=N, max_iter=max_iter)
    var = {"x": np.zeros((N + 1, nx)), "u": np.zeros((N, nu)), "lambda": np.zeros((N + 1, nx)),
           "s": np.ones((N, ng)), "nu": np.ones((N, ng))}  # Variable.reset(0, 0, 0, 1, 1)
    outs = [R.fmpc_solve(model, p, cfg, x0, var) for x0 in x0s]
    CASES[f"{name}/x0"] = np.array(x0s)
    CASES[f"{name}/N"] = np.array(N)
    CASES[f"{name}/max_iter"] = np.array(max_iter)
    for key in ("x", "u", "lambda", "s", "nu", "K", "kkt"):
        CASES[f"{name}/{key}"] = np.stack([o[key] for o in outs])
    CASES[f"{name}/n_trace"] = np.array([o["n_trace"] for o in outs])
    CASES[f"{name}/status"] = np.array([o["status"] for o in outs])


def main():
    swing = [0.0, np.pi, 0.0, 0.0]
    rnd = [list(v) for v in O.cartpole_x0(12, 0)]
    # BASELINE configs[1] shape (N=100, 10 iterations), reference termination and forced 10 iterations
    ddp_case("ddp_ref", [swing] + rnd, 100, max_iter=10)
    ddp_case("ddp_fixed", [swing] + rnd[:5], 100, max_iter=10, k_rel_norm_thre=0.0, cost_update_thre=0.0)
    ddp_case("ddp_reg2", rnd[:4], 60, max_iter=8, reg_type=2)
    ddp_case("ddp_default_500", rnd[:3], 100)
    # BASELINE configs[0]: TestDDPCartPole first tick (BoxQP branch), N=200 (rostest) and N=400 (literal)
    ddp_case("ddp_box200", [swing], 200, lo=[-15.0], hi=[15.0], max_iter=3, with_input_constraint=1)
    ddp_case("ddp_box400", [swing], 400, lo=[-15.0], hi=[15.0], max_iter=3, with_input_constraint=1)
    ddp_case("ddp_box_tight", rnd[:4], 100, lo=[-6.0], hi=[9.0], max_iter=12, with_input_constraint=1)
    # BASELINE configs[2] shape: FMPC cart-pole N=100, 10 iterations; TestFmpcOscillator first tick
    fmpc_case("fmpc_cartpole", "fmpc_cartpole", [swing] + rnd[:5], 100, 10)
    fmpc_case("fmpc_oscillator", "fmpc_oscillator", [[0.0, 1.0]], 400, 3)
    # TestBoxQP.cpp:39-98 through the reference's BoxQP<2> and BoxQP<Dynamic>
    H = np.array([[1.0, 0.0], [0.0, 0.5]])
    kats = [([1.5, 1.0], [-10, -10], [10, 10]), ([1.5, 1.0], [0.5, -2.0], [5.0, 2.0]),
            ([1.0, 1.5], [0.0, -1.0], [5.0, -0.5]), ([1.5, 1.0], [-5.0, -1.0], [-2.0, 2.0]),
            ([1.0, 1.5], [-5.0, -10.0], [-2.0, 10.0])]
    xs, rvs = [], []
    for dyn in (0, 1):
        for g, lo, hi in kats:
            x, rv = R.boxqp_solve2(H, g, lo, hi, dyn)
            xs.append(x)
            rvs.append(rv)
    CASES["boxqp/x"] = np.array(xs)
    CASES["boxqp/retval"] = np.array(rvs)
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **CASES)
    print("wrote", len(CASES), "arrays,", os.path.getsize(os.path.join(HERE, "reference_outputs.npz")), "bytes")


if __name__ == "__main__":
    main()
