"""BASELINE.json configs[3]: quadrotor iLQR, n_x=12, n_u=4, horizon 50, batch 8192, fp32 on the device.

The reference has neither this model nor a single-precision path: the fp64 oracle
(the restated DDPSolver around the same functor instantiated in double) is the yardstick, tolerance
rel. cost 1e-3 for fp32 and the fp64 tolerances of the cart-pole tests for the fp64
instantiation of the same kernels (this is what exercises the general n_u > 1 Cholesky path)."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
N = 50


def quadrotor_x0(B, seed):
    """p~U(-1,1)^3, rpy~U(-0.5,0.5)^3, v~U(-1,1)^3, w~U(-1,1)^3."""
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.uniform(-1, 1, (B, 3)), rng.uniform(-0.5, 0.5, (B, 3)), rng.uniform(-1, 1, (B, 3)),
                           rng.uniform(-1, 1, (B, 3))], axis=1)


def hover_inputs(B):
    u = np.zeros((B, N, 4))
    u[:, :, 0] = 9.80665  # m g with m = 1
    return u


def test_quadrotor_functor_derivatives_on_device(gpu):
    """The reference's derivative-check pattern (TestDDPCartPole.cpp:629-648: central differences, eps 1e-6,
    tol 1e-6) applied to the generated Jacobians, evaluated by the device functor in fp64."""
    rng = np.random.default_rng(0)
    x = quadrotor_x0(1, 1)[0]
    u = np.array([9.0, 0.1, -0.2, 0.05])
    d = gpu.model_eval("quadrotor_f64", 0.0, x[None], u[None])
    eps = 1e-6
    Fx, Fu = np.zeros((12, 12)), np.zeros((12, 4))
    for j in range(12):
        e = np.zeros(12)
        e[j] = eps
        Fx[:, j] = (gpu.model_eval("quadrotor_f64", 0.0, (x + e)[None], u[None])["x_next"][0]
                    - gpu.model_eval("quadrotor_f64", 0.0, (x - e)[None], u[None])["x_next"][0]) / (2 * eps)
    for j in range(4):
        e = np.zeros(4)
        e[j] = eps
        Fu[:, j] = (gpu.model_eval("quadrotor_f64", 0.0, x[None], (u + e)[None])["x_next"][0]
                    - gpu.model_eval("quadrotor_f64", 0.0, x[None], (u - e)[None])["x_next"][0]) / (2 * eps)
    assert np.linalg.norm(d["Fx"][0] - Fx) < 1e-6
    assert np.linalg.norm(d["Fu"][0] - Fu) < 1e-6
    o = O.model_eval("quadrotor", O.default_params("quadrotor"), 0.0, x, u)
    np.testing.assert_allclose(d["Fx"][0], o["Fx"], rtol=1e-12, atol=1e-14)
    del rng


@pytest.mark.parametrize("bwd_gs", ["16", "1"])
def test_quadrotor_fp64_parity(gpu, bwd_gs, monkeypatch):
    """Same kernels, double precision: cooperative (16 lanes / instance) and thread-per-instance K2."""
    monkeypatch.setenv("NMPC_B200_BWD_GS", bwd_gs)
    B = 96
    p = O.default_params("quadrotor")
    x0, u0 = quadrotor_x0(B, 4), hover_inputs(B)
    ref = O.ddp_solve_batch("quadrotor", p, O.ddp_config(max_iter=10, horizon_steps=N), x0, u0)
    solver = gpu.DDPSolver("quadrotor_f64", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps, c.max_iter = N, 10
    solver.solve_batch(0.0, x0, u0)
    assert np.array_equal(solver.iterations(), ref["iters"])
    assert np.array_equal(solver.status(), ref["status"])
    u = solver.controlData().u_list
    assert (np.max(np.abs(u - ref["u"]), axis=(1, 2)) / (1 + np.max(np.abs(ref["u"]), axis=(1, 2)))).max() <= 1e-8
    assert np.max(np.abs(solver.cost() - ref["cost"]) / np.abs(ref["cost"])) <= 1e-11


def test_quadrotor_fp32_config4(gpu):
    """configs[3] at full size: batch 8192, seed 4, fp32, 10 forced iterations; vs the fp64 oracle on a
    512-instance subset (the CPU oracle needs ~1 ms per instance)."""
    B, sub = 8192, 512
    p = O.default_params("quadrotor")
    x0, u0 = quadrotor_x0(B, 4), hover_inputs(B)
    solver = gpu.DDPSolver("quadrotor", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps, c.max_iter, c.k_rel_norm_thre, c.cost_update_thre = N, 10, 0.0, 0.0
    solver.solve_batch(0.0, x0, u0)
    cost = solver.cost()
    assert np.all(np.isfinite(cost)) and np.all(solver.status() >= 0)
    ref = O.ddp_solve_batch("quadrotor", p, O.ddp_config(max_iter=10, horizon_steps=N, k_rel_norm_thre=0.0,
                                                           cost_update_thre=0.0), x0[:sub], u0[:sub])
    rel = np.abs(cost[:sub] - ref["cost"]) / np.abs(ref["cost"])
    frac = float((rel <= 1e-3).mean())
    print(f"quadrotor fp32 vs fp64 oracle: rel cost error median {np.median(rel):.2e}, q99 {np.quantile(rel, 0.99):.2e}, "
          f"max {rel.max():.2e}; fraction within 1e-3: {frac:.4f}")
    # calibrated on B200: median 2e-7, q99 6e-7; ~0.2 % of the instances (large initial attitude, cost still
    # falling by 1-10 % per iteration at iteration 10) take a different backtracking step somewhere in fp32 and
    # end on a different -- equally valid -- iterate.  The gate is therefore a fraction, not a maximum.
    assert np.median(rel) <= 1e-5 and np.quantile(rel, 0.99) <= 1e-4
    assert frac >= 0.99
    # the optimiser must actually have optimised: cost far below the initial rollout's
    init = solver.trace()[:sub, 0, 1]
    assert np.all(cost[:sub] < init) and np.median(cost[:sub] / init) < 0.5


BOX_LO = np.array([7.0, -0.05, -0.05, -0.02])
BOX_HI = np.array([12.0, 0.05, 0.05, 0.02])
GOLDEN_BOX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_ddp_quadrotor.npz"))


def _rel_u(a, b):
    return np.max(np.abs(a - b), axis=(1, 2)) / (1 + np.max(np.abs(b), axis=(1, 2)))


def _oracle_box(max_iter, x0):
    g = GOLDEN_BOX
    return O.ddp_solve_batch("quadrotor", g["params"], O.ddp_config(max_iter=max_iter, horizon_steps=N,
                                                                     with_input_constraint=1), x0, g["u_init"],
                             u_lo=BOX_LO, u_hi=BOX_HI)


def _half_ulp(x0, seed):
    return x0 * (1.0 + 1.2e-16 * np.random.default_rng(seed).choice([-1, 0, 1], size=x0.shape))


@pytest.mark.parametrize("bwd_gs", ["1", "16"])
def test_quadrotor_input_limits_boxqp_nu4(gpu, bwd_gs, monkeypatch):
    """with_input_constraint for n_u = 4 (BoxQP with several free / clamped inputs per step, DDPSolver.hpp:450-497) on
    both fp64 K2 variants, and the fp32 column-split K2, against the reference headers' vectors
    (tests/golden/reference_ddp_quadrotor.npz).

    Two DDP iterations: parity at the fp64 tolerance.  Six iterations: the reference algorithm itself is discontinuous
    in its rounding noise by then (tests/test_ddp_quadrotor_box.py; tools/diag_boxqp_split.py traces one
    such step), so the CUDA path is held to the reference's OWN spread under a half-ulp change of x0, measured here with
    the oracle: no more instances off the reference's iterates, and no further off in cost, than twice that."""
    monkeypatch.setenv("NMPC_B200_BWD_GS", bwd_gs)
    B = 64
    g = GOLDEN_BOX
    p, x0, u0, lo, hi = g["params"], g["x0"], g["u_init"], BOX_LO, BOX_HI

    def solve(name, max_iter):
        s = gpu.DDPSolver(name, params=p, batch_capacity=B)
        c = s.config()
        c.horizon_steps, c.max_iter, c.with_input_constraint = N, max_iter, True
        s.setInputLimitsFunc((lo, hi))
        s.solve_batch(0.0, x0, u0)
        return s

    ref2 = _oracle_box(2, x0)
    np.testing.assert_array_equal(ref2["u"], g["it2/u"])
    # the limits bind (forwardPass itself does not clamp, DDPSolver.hpp:548 TODO: the feedback term may leave the box)
    assert np.any(np.isclose(ref2["u"][:, :, 1], hi[1])) and np.any(np.isclose(ref2["u"][:, :, 0], lo[0]))
    s = solve("quadrotor_f64", 2)
    assert np.array_equal(s.iterations(), ref2["iters"]) and np.array_equal(s.status(), ref2["status"])
    assert np.array_equal(s.n_forward(), ref2["n_fwd"]) and np.array_equal(s.n_backward(), ref2["n_bwd"])
    assert _rel_u(s.controlData().u_list, g["it2/u"]).max() <= 1e-8
    cost2 = g["it2/cost_list"].sum(axis=1)
    assert np.max(np.abs(s.cost() - cost2) / np.abs(cost2)) <= 1e-10
    k = s.k_list()
    assert (np.max(np.abs(k - ref2["k"]), axis=(1, 2)) / (1 + np.max(np.abs(ref2["k"]), axis=(1, 2)))).max() <= 1e-6

    ref6 = _oracle_box(6, x0)
    np.testing.assert_array_equal(ref6["u"], g["it6/u"])
    own = [_oracle_box(6, _half_ulp(x0, seed)) for seed in (1, 2, 3)]
    own_off = max(int((_rel_u(o["u"], ref6["u"]) > 1e-6).sum()) for o in own)
    own_cost = max(float(np.max(np.abs(o["cost"] - ref6["cost"]) / np.abs(ref6["cost"]))) for o in own)
    s = solve("quadrotor_f64", 6)
    rel_c = np.abs(s.cost() - ref6["cost"]) / np.abs(ref6["cost"])
    off = int((_rel_u(s.controlData().u_list, ref6["u"]) > 1e-6).sum())
    assert np.array_equal(s.iterations(), ref6["iters"]) and np.array_equal(s.status(), ref6["status"])
    assert off <= 2 * own_off and rel_c.max() <= 2 * own_cost, (off, own_off, rel_c.max(), own_cost)
    assert np.median(rel_c) <= 1e-2 and rel_c.max() <= 0.1, (np.median(rel_c), rel_c.max())
    assert np.all(s.cost() < s.trace()[:, 0, 1])  # every instance improved on its initial rollout
    monkeypatch.delenv("NMPC_B200_BWD_GS")

    s32 = solve("quadrotor", 6)  # fp32, column-split K2 with 12 warps per tile
    assert np.all(np.isfinite(s32.controlData().u_list)) and np.all(s32.status() >= 0)
    rel = np.abs(s32.cost() - ref6["cost"]) / np.abs(ref6["cost"])
    assert np.median(rel) <= 2e-2 and rel.max() <= 0.2, (np.median(rel), rel.max())
    assert np.all(s32.cost() < s32.trace()[:, 0, 1])
