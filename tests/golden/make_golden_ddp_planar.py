#!/usr/bin/env python
"""Golden vectors for the control-limited DDP backward pass with MORE THAN ONE input, produced by the REFERENCE's
unmodified DDPSolver.h/.hpp + BoxQP.h (oracle/_ref, compiled against oracle/ref/eigen_shim): the planar quadrotor as a
DDPProblem<6, 2>, so that BoxQP<2> runs with both, one or no input clamped (DDPSolver.hpp:450-497, BoxQP.h:84-214).
The reference's own tests only reach the limited backward pass with one input (TestDDPCartPole.cpp).

    python tests/golden/make_golden_ddp_planar.py     ->  tests/golden/reference_ddp_planar.npz
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
    return np.concatenate([rng.uniform(-1.5, 1.5, (n, 2)), rng.uniform(-0.5, 0.5, (n, 1)), rng.uniform(-0.8, 0.8, (n, 3))],
                          axis=1)


def case(name, params, x0s, N, lo=None, hi=None, u_init=None, **cfg_kw):
    cfg = R.ddp_config(horizon_steps=N, **cfg_kw)
    u_init = np.zeros((N, 2)) if u_init is None else np.asarray(u_init)
    outs = [R.ddp_solve_planar(params, cfg, x0, u_init, u_lo=lo, u_hi=hi) for x0 in x0s]
    CASES[f"{name}/params"] = np.asarray(params)
    CASES[f"{name}/x0"] = np.array(x0s)
    CASES[f"{name}/u_init"] = u_init
    CASES[f"{name}/N"] = np.array(N)
    for k, v in cfg_kw.items():
        CASES[f"{name}/cfg_{k}"] = np.array(v)
    if lo is not None:
        CASES[f"{name}/limits"] = np.array([lo, hi], dtype=np.float64)
    for key in ("x", "u", "cost_list", "trace"):
        CASES[f"{name}/{key}"] = np.stack([o[key] for o in outs])
    CASES[f"{name}/n_trace"] = np.array([o["n_trace"] for o in outs])
    CASES[f"{name}/solve_ret"] = np.array([o["solve_ret"] for o in outs])
    u = CASES[f"{name}/u"]
    clamp = "-" if lo is None else f"at lo {np.mean(u <= np.asarray(lo) + 1e-12):.2f} at hi {np.mean(u >= np.asarray(hi) - 1e-12):.2f}"
    print(name, "iters", CASES[f"{name}/n_trace"] - 1, "ret", CASES[f"{name}/solve_ret"], "clamped", clamp)


def main():
    p = O.default_params("planar_quadrotor")
    hover = 0.5 * p[1] * 9.80665
    N = 60
    x0s = list(planar_x0(4, 3))
    uh = np.full((N, 2), hover)
    # no limits: the plain n_u = 2 backward pass (LLT of a 2 x 2 Quu)
    case("planar_free", p, x0s, N, u_init=uh, max_iter=60)
    # generous limits: the box QP runs but seldom clamps
    case("planar_box_wide", p, x0s, N, lo=[0.0, 0.0], hi=[2.0 * hover, 2.0 * hover], u_init=uh, with_input_constraint=1,
         max_iter=60)
    # tight symmetric limits: both inputs saturate over long stretches
    case("planar_box_tight", p, x0s, N, lo=[0.8 * hover, 0.8 * hover], hi=[1.2 * hover, 1.2 * hover], u_init=uh,
         with_input_constraint=1, max_iter=60)
    # different limits per input: one rotor saturates while the other stays free (the mixed clamped/free Hessian block)
    case("planar_box_mixed", p, x0s, N, lo=[0.9 * hover, 0.2 * hover], hi=[1.05 * hover, 3.0 * hover], u_init=uh,
         with_input_constraint=1, max_iter=60)
    # coupled inputs (cross weight in Luu), fixed iteration count, second regularisation type
    q = p.copy()
    q[12] = 0.4 * q[11]
    case("planar_box_cross_fixed", q, x0s, N, lo=[0.85 * hover, 0.85 * hover], hi=[1.3 * hover, 1.1 * hover], u_init=uh,
         with_input_constraint=1, max_iter=8, k_rel_norm_thre=0.0, cost_update_thre=0.0, cost_update_ratio_thre=0.0,
         lambda_thre=0.0, reg_type=2)
    # zero initial inputs (outside the box): the first forward pass starts from an infeasible input sequence
    case("planar_box_cold", p, x0s, N, lo=[0.7 * hover, 0.7 * hover], hi=[1.4 * hover, 1.4 * hover],
         with_input_constraint=1, max_iter=80)
    # The two cases above are ones where the REFERENCE itself is discontinuous in the last bit of its input from the
    # third iteration on (tests/test_ddp_planar.py::test_reference_splits_on_the_last_bit_at_a_limit); their first two
    # iterations are not, and are what an implementation with different rounding can be held to exactly.
    case("planar_box_cross_fixed_it2", q, x0s, N, lo=[0.85 * hover, 0.85 * hover], hi=[1.3 * hover, 1.1 * hover],
         u_init=uh, with_input_constraint=1, max_iter=2, k_rel_norm_thre=0.0, cost_update_thre=0.0,
         cost_update_ratio_thre=0.0, lambda_thre=0.0, reg_type=2)
    case("planar_box_cold_it2", p, x0s, N, lo=[0.7 * hover, 0.7 * hover], hi=[1.4 * hover, 1.4 * hover],
         with_input_constraint=1, max_iter=2)
    out = os.path.join(HERE, "reference_ddp_planar.npz")
    np.savez_compressed(out, **CASES)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
