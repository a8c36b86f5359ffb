#!/usr/bin/env python
"""Golden vectors for BoxQP<4> inside the control-limited DDP backward pass: the REFERENCE's unmodified DDPSolver<12, 4>
(oracle/_ref, Eigen shim) on the 3-D quadrotor functor (include/nmpc_b200/models/quadrotor.h behind the reference's
DDPProblem interface, oracle/ref/ref_ddp.cpp), with the input limits of tests/test_quadrotor_gpu.py, after 2 and after
6 iterations.

    python tests/golden/make_golden_ddp_quadrotor.py     ->  tests/golden/reference_ddp_quadrotor.npz
Needs /root/reference (this container only); the committed .npz is what the tests read.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
import ref_lib as R  # noqa: E402
from test_quadrotor_gpu import N, hover_inputs, quadrotor_x0  # noqa: E402

B = 64
LO = np.array([7.0, -0.05, -0.05, -0.02])
HI = np.array([12.0, 0.05, 0.05, 0.02])


def main():
    p = O.default_params("quadrotor")
    x0, u0 = quadrotor_x0(B, 9), hover_inputs(B)
    out = {"params": p, "x0": x0, "u_init": u0, "lo": LO, "hi": HI, "N": np.array(N)}
    for mi in (2, 6):
        cfg = R.ddp_config(max_iter=mi, horizon_steps=N, with_input_constraint=1)
        outs = [R.ddp_solve_quadrotor(p, cfg, x0[b], u0[b], u_lo=LO, u_hi=HI) for b in range(B)]
        out[f"it{mi}/u"] = np.stack([o["u"] for o in outs])
        out[f"it{mi}/cost_list"] = np.stack([o["cost_list"] for o in outs])
        out[f"it{mi}/trace"] = np.stack([o["trace"] for o in outs])
        out[f"it{mi}/n_trace"] = np.array([o["n_trace"] for o in outs])
        out[f"it{mi}/solve_ret"] = np.array([o["solve_ret"] for o in outs])
        print(mi, "iters", np.bincount(out[f"it{mi}/n_trace"] - 1), "at limit", np.mean((out[f"it{mi}/u"] <= LO) | (out[f"it{mi}/u"] >= HI)))
    path = os.path.join(HERE, "reference_ddp_quadrotor.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
