#!/usr/bin/env python
"""Small solves that walk every hand-synchronised kernel once, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_cases.py [case ...]

Cases (each a few instances, short horizons, 2-3 iterations -- the tools slow kernels down 10-100x):
  lanes      K1+K2 with four lanes per instance (producer warps, mbarrier ring, smem exchange) + split line search
  retry      the lambda-retry path of the same kernel (indefinite Quu) and BoxQP limits
  tiny       horizons 1, 2, 3, 7 (shorter than the rings) and a ragged batch
  variants   the thread-per-instance fused K2, the phased line search, the shuffle exchange, the persistent tile kernel
  fmpc       F1..F4 with loader warps (cart-pole), n_u = 2 (planar quadrotor), windowed inequality dimension
  quad       quadrotor fp32: column-split K2 over 12 warps (TMA tile ring)
  wide       centroidal motion: 16 lanes per instance, cp.async ring
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nmpc_b200  # noqa: E402
import oracle_lib as O  # noqa: E402  (inputs only)


def ddp(model, B, N, x0, u0, env=None, **cfg):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        s = nmpc_b200.DDPSolver(model, batch_capacity=B, params=cfg.pop("params", None))
        c = s.config()
        c.horizon_steps = N
        limits = cfg.pop("limits", None)
        for k, v in cfg.items():
            setattr(c, k, v)
        if limits is not None:
            s.setInputLimitsFunc(limits)
        s.solve_batch(0.0, x0, u0)
        assert np.isfinite(s.cost()).all()
        s.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def case_lanes():
    B, N = 37, 24
    ddp("cartpole", B, N, O.cartpole_x0(B, 1), np.zeros((B, N, 1)), max_iter=3)
    ddp("cartpole", B, N, O.cartpole_x0(B, 1), np.zeros((B, N, 1)), max_iter=3, k_rel_norm_thre=0.0, cost_update_thre=0.0)
    ddp("bipedal", 5, 20, np.zeros((5, 2)), np.zeros((5, 20, 1)), max_iter=2)


def case_retry():
    B, N = 33, 20
    p = O.default_params("cartpole")
    p[8] = -5e-4  # running_u < 0: Quu indefinite, lambda grows inside the kernel
    ddp("cartpole", B, N, O.cartpole_x0(B, 21), np.zeros((B, N, 1)), max_iter=3, params=p)
    ddp("cartpole", B, N, O.cartpole_x0(B, 12), np.zeros((B, N, 1)), max_iter=3, with_input_constraint=True,
        limits=(np.array([-6.0]), np.array([9.0])))


def case_tiny():
    for N in (1, 2, 3, 7):
        ddp("cartpole", 5, N, O.cartpole_x0(5, 40 + N), np.zeros((5, N, 1)), max_iter=3)


def case_variants():
    B, N = 40, 16
    x0, u0 = O.cartpole_x0(B, 3), np.zeros((B, N, 1))
    for env in ({"NMPC_B200_BWD_LANES": "0"}, {"NMPC_B200_FWD_SPLIT": "0"}, {"NMPC_B200_BWD_LANES": "2"},
                {"NMPC_B200_TILE": "1"}, {"NMPC_B200_BWD_FUSED": "0"}):
        ddp("cartpole", B, N, x0, u0, env=env, max_iter=3, k_rel_norm_thre=0.0, cost_update_thre=0.0)


def fmpc(model, B, N, x0, u_init=0.0, **cfg):
    s = nmpc_b200.FmpcSolver(model, batch_capacity=B)
    c = s.config()
    c.horizon_steps, c.max_iter = N, 2
    for k, v in cfg.items():
        setattr(c, k, v)
    v = s.make_variable(B)
    v.reset(0.0, u_init, 0.0, 1.0, 1.0)
    s.solve_batch(0.0, x0, v)
    s.close()


def case_fmpc():
    fmpc("cartpole", 35, 20, O.cartpole_x0(35, 3))
    fmpc("cartpole", 9, 3, O.cartpole_x0(9, 4), enable_line_search=True)
    rng = np.random.default_rng(0)
    fmpc("planar_quadrotor", 33, 12, rng.uniform(-0.5, 0.5, (33, 6)), u_init=4.9)
    fmpc("cartpole_windowed", 20, 40, O.cartpole_x0(20, 5))


def case_quad():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_quadrotor_gpu import N, hover_inputs, quadrotor_x0

    B = 40
    ddp("quadrotor", B, N, quadrotor_x0(B, 1), hover_inputs(B), max_iter=2)


def case_wide():
    """Centroidal motion (n_x = 9, n_u = 16 / 0): the wide K2 and, with input limits, the cooperative K2 + BoxQP."""
    from test_centroidal_motion import N, X0

    B = 3
    x0 = np.repeat(X0, B, axis=0)
    ddp("centroidal_motion", B, N, x0, np.zeros((B, N, 16)), max_iter=2)
    ddp("centroidal_motion", B, N, x0, np.zeros((B, N, 16)), max_iter=2, with_input_constraint=True,
        limits=(np.zeros(16), np.full(16, 200.0)))


CASES = {"lanes": case_lanes, "retry": case_retry, "tiny": case_tiny, "variants": case_variants, "fmpc": case_fmpc,
         "quad": case_quad, "wide": case_wide}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        CASES[n]()
        print("case", n, "done", flush=True)
