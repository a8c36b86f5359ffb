#!/usr/bin/env python
"""Golden vectors for the two FMPC code paths the reference's own tests never reach, produced by the REFERENCE's
unmodified FmpcSolver.h/.hpp (oracle/_ref, compiled against oracle/ref/eigen_shim):

  * n_u = 2 (planar quadrotor, FmpcSolver<6, 2, 4>): Eigen::LDLT of G with diagonal pivoting, and the
    Eigen::FullPivLU fallback / break_if_llt_fails exit when LDLT reports NumericalIssue (FmpcSolver.hpp:596-617);
  * a time-varying inequality dimension (windowed cart-pole, FmpcSolver<4, 1, Eigen::Dynamic>, FmpcProblem.h:62-86).

    python tests/golden/make_golden_fmpc_extra.py     ->  tests/golden/reference_fmpc_extra.npz
Needs /root/reference (this container only); the committed .npz is what the tests read.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
import ref_lib as R  # noqa: E402

CASES = {}


def planar_x0(n, seed):
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.uniform(-1, 1, (n, 2)), rng.uniform(-0.4, 0.4, (n, 1)), rng.uniform(-0.5, 0.5, (n, 2)),
                           rng.uniform(-0.5, 0.5, (n, 1))], axis=1)


def run(name, model, params, cfg_kw, x0s, var, t0=0.0):
    nx, nu, ng, _ = O.model_dims(model)
    cfg = O.fmpc_config(**cfg_kw)
    outs = [R.fmpc_solve(model, params, cfg, x0, var, t0=t0) for x0 in x0s]
    CASES[f"{name}/params"] = np.asarray(params)
    CASES[f"{name}/x0"] = np.array(x0s)
    CASES[f"{name}/t0"] = np.array(t0)
    for k, v in cfg_kw.items():
        CASES[f"{name}/cfg_{k}"] = np.array(v)
    for k in ("x", "u", "lambda", "s", "nu"):
        CASES[f"{name}/var_{k}"] = var[k]
    for key in ("x", "u", "lambda", "s", "nu", "K", "kkt"):
        CASES[f"{name}/{key}"] = np.stack([o[key] for o in outs])
    CASES[f"{name}/n_trace"] = np.array([o["n_trace"] for o in outs])
    CASES[f"{name}/status"] = np.array([o["status"] for o in outs])
    print(name, "status", CASES[f"{name}/status"], "kkt[0]", outs[0]["kkt"][:outs[0]["n_trace"]])


def main():
    # ---- planar quadrotor, regular (positive definite G: pivoted LDLT with a non-trivial pivot order)
    p = O.default_params("fmpc_planar_quadrotor")
    N = 60
    hover = 0.5 * p[1] * 9.80665
    var = {"x": np.zeros((N + 1, 6)), "u": np.full((N, 2), hover), "lambda": np.zeros((N + 1, 6)),
           "s": np.ones((N, 4)), "nu": np.ones((N, 4))}
    for it in (1, 2, 3, 5, 10):
        run(f"planar_it{it}", "fmpc_planar_quadrotor", p, dict(horizon_steps=N, max_iter=it), list(planar_x0(4, 1)), var)
    # ---- planar quadrotor, G = dt [[0, c], [c, 0]] at the last step: LDLT reports NumericalIssue
    q = p.copy()
    q[11], q[12] = 0.0, 0.05  # running_u = 0, cross weight only
    q[13:19] = 0.0  # no terminal cost: P_N = 0
    N2 = 20
    var0 = {"x": np.zeros((N2 + 1, 6)), "u": np.full((N2, 2), hover), "lambda": np.zeros((N2 + 1, 6)),
            "s": np.ones((N2, 4)), "nu": np.zeros((N2, 4))}  # nu = 0: no D^T diag(nu / s) D on the diagonal of G
    run("planar_fullpivlu", "fmpc_planar_quadrotor", q, dict(horizon_steps=N2, max_iter=2), list(planar_x0(3, 2)), var0)
    run("planar_break", "fmpc_planar_quadrotor", q, dict(horizon_steps=N2, max_iter=2, break_if_llt_fails=1),
        list(planar_x0(2, 2)), var0)
    # ---- windowed cart-pole: inequality dimension 4 inside [0.3, 0.7), 2 outside
    pw = O.default_params("fmpc_cartpole_windowed")
    Nw = 100
    varw = {"x": np.zeros((Nw + 1, 4)), "u": np.zeros((Nw, 1)), "lambda": np.zeros((Nw + 1, 4)),
            "s": np.ones((Nw, 4)), "nu": np.ones((Nw, 4))}
    x0w = [[0.0, np.pi, 0.0, 0.0]] + [list(v) for v in O.cartpole_x0(3, 5)]
    for it in (1, 3, 10):
        run(f"windowed_it{it}", "fmpc_cartpole_windowed", pw, dict(horizon_steps=Nw, max_iter=it), x0w, varw)
    run("windowed_t0", "fmpc_cartpole_windowed", pw, dict(horizon_steps=Nw, max_iter=5), x0w[:2], varw, t0=0.255)
    run("windowed_initcomp", "fmpc_cartpole_windowed", pw,
        dict(horizon_steps=Nw, max_iter=3, init_complementary_variable=1), x0w[:2], varw)
    run("windowed_linesearch", "fmpc_cartpole_windowed", pw, dict(horizon_steps=Nw, max_iter=3, enable_line_search=1),
        x0w[:2], varw)
    np.savez_compressed(os.path.join(HERE, "reference_fmpc_extra.npz"), **CASES)
    print("wrote", len(CASES), "arrays")


if __name__ == "__main__":
    main()
