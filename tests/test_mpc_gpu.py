"""GPU parity of the device-resident receding-horizon loops (nmpc_b200_ddp_run_mpc / _fmpc_run_mpc) against the
same loops driven from the host around the CPU oracle, tick by tick.

The loop bodies are the reference's own callers of the hot path: TestDDPBipedal.cpp:243-268 (plant = model
prediction, shifted warm start), TestDDPCartPole.cpp:313-343 + :388-396 (simulated plant with sim_dt, clamped
input, unshifted warm start, BoxQP-constrained solves with max_iter = 3)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _ref_zmp(tt):
    tt += 1e-6  # TestDDPBipedal.cpp:203-204
    if tt <= 1.5 or tt >= 20.0 - 1.5:
        return 0.0
    return 0.15 if int(np.floor((tt - 1.0) / 1.0)) % 2 == 0 else -0.15


def test_bipedal_mpc_loop_matches_host_loop(gpu):
    """TestDDPBipedal.TestCase1's loop, first 1.6 s (it crosses the first ZMP switch at t = 1.5 s): instance 0 starts
    from the test's state (0, 0) and must meet its per-tick threshold |planned_zmp - ref_zmp| < 1e-2; the other
    instances start from perturbed CoM states.  Every tick is compared with the oracle-driven host loop."""
    p = O.default_params("bipedal")
    N, B, ticks = 300, 6, 160
    dt = p[0]
    rng = np.random.default_rng(11)
    x0 = np.concatenate([np.zeros((1, 2)), rng.uniform(-0.02, 0.02, (B - 1, 2))])
    u0 = np.zeros((B, N, 1))

    solver = gpu.DDPSolver("bipedal", params=p, batch_capacity=B)
    solver.config().horizon_steps = N
    got = solver.run_mpc(0.0, x0, u0, n_ticks=ticks, tick_dt=dt, plant="model", shift_inputs=True)

    cfg = O.ddp_config(horizon_steps=N)
    t, x, u = 0.0, x0.copy(), u0.copy()
    for k in range(ticks):
        r = O.ddp_solve_batch("bipedal", p, cfg, x, u, t0=t)
        np.testing.assert_allclose(got["x"][:, k], x, rtol=0, atol=1e-9, err_msg=f"tick {k}")
        np.testing.assert_allclose(got["u"][:, k], r["u"][:, 0], rtol=0, atol=1e-8, err_msg=f"tick {k}")
        assert np.array_equal(got["iters"][:, k], r["iters"]), f"tick {k}"
        assert np.array_equal(got["status"][:, k], r["status"]), f"tick {k}"
        assert abs(got["u"][0, k, 0] - _ref_zmp(t)) < 1e-2  # the reference's own check (TestDDPBipedal.cpp:256)
        t = (k + 1) * dt
        x = r["x"][:, 1].copy()
        u = np.concatenate([r["u"][:, 1:], r["u"][:, -1:]], axis=1)
    np.testing.assert_allclose(got["x"][:, ticks], x, rtol=0, atol=1e-9)
    # the handle holds the last solve
    np.testing.assert_allclose(solver.controlData().u_list, r["u"], rtol=0, atol=1e-8)


def test_cartpole_mpc_loop_simulated_plant(gpu):
    """TestDDPCartPole's loop: horizon 2 s / 0.01 s = 200 steps, max_iter 3, BoxQP input limits +-15 N, MPC tick
    4 ms, plant integrated at 2 ms, applied input clamped, warm start = previous u_list unshifted."""
    p = O.default_params("cartpole")
    N, B, ticks = 200, 5, 40
    mpc_dt, sim_dt = 0.004, 0.002
    x0 = np.concatenate([[[0.0, np.pi, 0.0, 0.0]], O.cartpole_x0(B - 1, 5)])
    u0 = np.zeros((B, N, 1))
    lo, hi = np.array([-15.0]), np.array([15.0])

    solver = gpu.DDPSolver("cartpole", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps, c.max_iter, c.with_input_constraint = N, 3, True
    solver.setInputLimitsFunc((lo, hi))
    got = solver.run_mpc(0.0, x0, u0, n_ticks=ticks, tick_dt=mpc_dt, plant="sim", shift_inputs=False, clamp_u0=True,
                         sim_dt=sim_dt, n_substeps=2)

    cfg = O.ddp_config(horizon_steps=N, max_iter=3, with_input_constraint=1)
    p_sim = p.copy()
    p_sim[0] = sim_dt  # the oracle's stateEq with dt = sim_dt is the plant (TestDDPCartPole.cpp:330)
    t, x, u = 0.0, x0.copy(), u0.copy()
    for k in range(ticks):
        r = O.ddp_solve_batch("cartpole", p, cfg, x, u, t0=t, u_lo=lo, u_hi=hi)
        ua = np.clip(r["u"][:, 0], lo, hi)
        np.testing.assert_allclose(got["x"][:, k], x, rtol=0, atol=1e-8, err_msg=f"tick {k}")
        np.testing.assert_allclose(got["u"][:, k], ua, rtol=0, atol=1e-7, err_msg=f"tick {k}")
        assert np.array_equal(got["iters"][:, k], r["iters"]), f"tick {k}"
        for _ in range(2):
            x = np.stack([O.model_eval("cartpole", p_sim, t, x[b], ua[b])["x_next"] for b in range(B)])
        t = (k + 1) * mpc_dt
        u = r["u"].copy()
    np.testing.assert_allclose(got["x"][:, ticks], x, rtol=0, atol=1e-8)


def test_mpc_argument_errors(gpu):
    solver = gpu.DDPSolver("bipedal", batch_capacity=2)
    solver.config().horizon_steps = 20
    x0, u0 = np.zeros((2, 2)), np.zeros((2, 20, 1))
    with pytest.raises(gpu.NmpcB200Error) as e:  # the bipedal functor has no stateEq(t, x, u, dt)
        solver.run_mpc(0.0, x0, u0, n_ticks=2, tick_dt=0.01, plant="sim", sim_dt=0.005)
    assert e.value.code == 7
    with pytest.raises(gpu.NmpcB200Error):
        solver.run_mpc(0.0, x0, u0, n_ticks=0, tick_dt=0.01)
    with pytest.raises(gpu.NmpcB200Error):  # clamp without limits
        solver.run_mpc(0.0, x0, u0, n_ticks=2, tick_dt=0.01, clamp_u0=True)
    with pytest.raises(ValueError):  # initial_u_list length (DDPSolver.hpp:41-45)
        solver.run_mpc(0.0, x0, np.zeros((2, 19, 1)), n_ticks=2, tick_dt=0.01)
