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


def test_cartpole_mpc_loop_with_limits_that_depend_on_time(gpu):
    """The same loop with input limits that are a genuine function of time (DDPSolver.h:282-285): the reference
    evaluates input_limits_func_(t_i) at every solve (DDPSolver.hpp:470), so every tick of the device-resident loop has
    its own limit table (nmpc_b200_ddp_set_input_limits_mpc); the clamp of the applied input uses the tick's time."""
    p = O.default_params("cartpole")
    N, B, ticks = 100, 4, 25
    mpc_dt, sim_dt, dt = 0.004, 0.002, p[0]
    x0 = np.concatenate([[[0.0, np.pi, 0.0, 0.0]], O.cartpole_x0(B - 1, 8)])
    u0 = np.zeros((B, N, 1))

    def limits(t):
        w = 15.0 - 10.0 * min(max(t / 0.6, 0.0), 1.0)  # the band closes from +-15 N to (-5, 2.5) N within 0.6 s
        return np.array([-w]), np.array([0.5 * w])

    solver = gpu.DDPSolver("cartpole", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps, c.max_iter, c.with_input_constraint = N, 3, True
    solver.setInputLimitsFunc(limits)
    got = solver.run_mpc(0.0, x0, u0, n_ticks=ticks, tick_dt=mpc_dt, plant="sim", shift_inputs=False, clamp_u0=True,
                         sim_dt=sim_dt, n_substeps=2)

    cfg = O.ddp_config(horizon_steps=N, max_iter=3, with_input_constraint=1)
    p_sim = p.copy()
    p_sim[0] = sim_dt
    t, x, u = 0.0, x0.copy(), u0.copy()
    clamped = 0
    for k in range(ticks):
        lo = np.array([limits(t + i * dt)[0] for i in range(N)])
        hi = np.array([limits(t + i * dt)[1] for i in range(N)])
        r = O.ddp_solve_cartpole_tv_limits(p, cfg, x, u, lo, hi, t0=t)
        ua = np.clip(r["u"][:, 0], lo[0], hi[0])
        clamped += int(np.sum(ua != r["u"][:, 0]))
        np.testing.assert_allclose(got["x"][:, k], x, rtol=0, atol=1e-8, err_msg=f"tick {k}")
        np.testing.assert_allclose(got["u"][:, k], ua, rtol=0, atol=1e-7, err_msg=f"tick {k}")
        assert np.array_equal(got["iters"][:, k], r["iters"]), f"tick {k}"
        for _ in range(2):
            x = np.stack([O.model_eval("cartpole", p_sim, t, x[b], ua[b])["x_next"] for b in range(B)])
        t = (k + 1) * mpc_dt
        u = r["u"].copy()
    np.testing.assert_allclose(got["x"][:, ticks], x, rtol=0, atol=1e-8)
    # without a per-tick table (the raw C entry point with limits that vary along the horizon) the loop still refuses
    lo0 = np.array([limits(i * dt)[0] for i in range(N)])
    hi0 = np.array([limits(i * dt)[1] for i in range(N)])
    import ctypes as C

    from nmpc_b200._capi import check, lib
    check(lib().nmpc_b200_ddp_set_input_limits_horizon(solver._h, N, lo0.ctypes.data_as(C.c_void_p),
                                                       hi0.ctypes.data_as(C.c_void_p)))
    solver._limits_func = None
    with pytest.raises(gpu.NmpcB200Error) as e:
        solver.run_mpc(0.0, x0, u0, n_ticks=ticks + 5, tick_dt=mpc_dt, plant="sim", shift_inputs=False, clamp_u0=True,
                       sim_dt=sim_dt, n_substeps=2)
    assert e.value.code == 7


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


def _fmpc_var(solver, B):
    v = solver.make_variable(B)
    v.reset(0.0, 0.0, 0.0, 1.0, 1.0)  # TestFmpcOscillator.cpp:153-154, TestFmpcCartPole.cpp:329-330
    return v


def _var_dict(v):
    return {"x": v.x_list, "u": v.u_list, "lambda": v.lambda_list, "s": v.s_list, "nu": v.nu_list}


def test_fmpc_oscillator_mpc_loop(gpu):
    """TestFmpcOscillator.SolveMpc's loop (horizon 4 s / 0.01 s, max_iter 3, sim_dt 5 ms), first 0.6 s, for a small
    batch of initial states; instance 0 is the test's own (0, 1).  Every tick: status Succeeded or
    MaxIterationReached (:170), constraints hold (:180), u_list[0] and current_x track the host loop around the oracle."""
    N, B, ticks, sim_dt = 400, 4, 120, 0.005
    p = O.default_params("fmpc_oscillator")
    x0 = np.array([[0.0, 1.0], [0.1, 0.8], [-0.1, 0.9], [0.05, 1.1]])
    solver = gpu.FmpcSolver("oscillator", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps, c.max_iter = N, 3
    got = solver.run_mpc(0.0, x0, _fmpc_var(solver, B), n_ticks=ticks, tick_dt=sim_dt, plant="sim", sim_dt=sim_dt)
    assert set(np.unique(got["status"])) <= {1, 5}

    ocfg = O.fmpc_config(horizon_steps=N, max_iter=3)
    ovar = _var_dict(_fmpc_var(solver, B))
    p_sim = p.copy()
    p_sim[0] = sim_dt
    x, t = x0.copy(), 0.0
    for k in range(ticks):
        r = O.fmpc_solve_batch("fmpc_oscillator", p, ocfg, x, ovar, t0=t)
        ovar = {key: r[key] for key in ("x", "u", "lambda", "s", "nu")}
        u = r["u"][:, 0]
        np.testing.assert_allclose(got["x"][:, k], x, rtol=0, atol=1e-7, err_msg=f"tick {k}")
        np.testing.assert_allclose(got["u"][:, k], u, rtol=0, atol=1e-6, err_msg=f"tick {k}")
        assert np.array_equal(got["status"][:, k], r["status"]), f"tick {k}"
        kkt = np.array([r["trace"][b, r["n_trace"][b] - 1, 1] for b in range(B)])
        np.testing.assert_allclose(got["kkt_error"][:, k], kkt, rtol=1e-5, atol=1e-9, err_msg=f"tick {k}")
        g = np.stack([-got["x"][:, k, 1] - 0.05, -got["u"][:, k, 0] - 1.0, got["u"][:, k, 0] - 0.9], axis=1)
        assert np.all(g <= 0), (k, g)
        x = np.stack([O.model_eval("fmpc_oscillator", p_sim, t, x[b], u[b])["x_next"] for b in range(B)])
        t = (k + 1) * sim_dt
    np.testing.assert_allclose(got["x"][:, ticks], x, rtol=0, atol=1e-7)
    # the handle holds the last solve's Variable
    np.testing.assert_allclose(solver.variable().u_list, r["u"], rtol=0, atol=1e-6)


def test_fmpc_cartpole_mpc_loop_with_feedback(gpu):
    """TestFmpcCartPole's loop: horizon 2 s / 0.01 s, max_iter 5, MPC tick 4 ms, plant at 2 ms with the feedback term
    u_list[0] + K_0 (x_list[0] - current_x) at every sub-step (:351-356)."""
    N, B, ticks = 200, 3, 25
    mpc_dt, sim_dt = 0.004, 0.002
    p = O.default_params("fmpc_cartpole")
    x0 = np.array([[0.0, 0.3, 0.0, 0.0], [0.2, -0.2, 0.1, 0.0], [-0.3, 0.1, 0.0, 0.2]])
    solver = gpu.FmpcSolver("cartpole", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps, c.max_iter = N, 5
    got = solver.run_mpc(0.0, x0, _fmpc_var(solver, B), n_ticks=ticks, tick_dt=mpc_dt, plant="sim", sim_dt=sim_dt,
                         n_substeps=2, feedback=True)

    ocfg = O.fmpc_config(horizon_steps=N, max_iter=5)
    ovar = _var_dict(_fmpc_var(solver, B))
    p_sim = p.copy()
    p_sim[0] = sim_dt
    x, t = x0.copy(), 0.0
    for k in range(ticks):
        r = O.fmpc_solve_batch("fmpc_cartpole", p, ocfg, x, ovar, t0=t)
        ovar = {key: r[key] for key in ("x", "u", "lambda", "s", "nu")}
        np.testing.assert_allclose(got["x"][:, k], x, rtol=0, atol=1e-6, err_msg=f"tick {k}")
        np.testing.assert_allclose(got["u"][:, k], r["u"][:, 0], rtol=1e-6, atol=1e-6, err_msg=f"tick {k}")
        assert np.array_equal(got["status"][:, k], r["status"]), f"tick {k}"
        K0 = r["K"][:, 0].reshape(B, 4, 1).transpose(0, 2, 1)  # column-major NU x NX
        for _ in range(2):
            u = r["u"][:, 0] + np.einsum("bij,bj->bi", K0, r["x"][:, 0] - x)
            x = np.stack([O.model_eval("fmpc_cartpole", p_sim, t, x[b], u[b])["x_next"] for b in range(B)])
        t = (k + 1) * mpc_dt
    np.testing.assert_allclose(got["x"][:, ticks], x, rtol=0, atol=1e-6)


def test_fmpc_mpc_argument_errors(gpu):
    solver = gpu.FmpcSolver("oscillator", batch_capacity=2)
    solver.config().horizon_steps = 20
    x0 = np.array([[0.0, 1.0], [0.0, 0.9]])
    with pytest.raises(gpu.NmpcB200Error):
        solver.run_mpc(0.0, x0, _fmpc_var(solver, 2), n_ticks=0, tick_dt=0.005)
    bad = _fmpc_var(solver, 2)
    bad.s_list[1, 3, 0] = -1.0  # checkVariable (FmpcSolver.hpp:348-361)
    with pytest.raises(gpu.NmpcB200Error) as e:
        solver.run_mpc(0.0, x0, bad, n_ticks=2, tick_dt=0.005)
    assert "non-negative" in str(e.value)


def test_trace_dump_files_read_like_the_reference_plot_scripts(gpu, tmp_path):
    """dumpTraceDataList() writes the tables nmpc_ddp/scripts/plotDDPTraceData.py:9-10 and its FMPC twin load with
    np.genfromtxt(..., delimiter=' ', names=True): same header names, one row per trace entry (DDPSolver.hpp:563-598,
    FmpcSolver.hpp:260-283)."""
    ddp = gpu.DDPSolver("cartpole", batch_capacity=2)
    ddp.config().max_iter = 10
    ddp.solve_batch(0.0, O.cartpole_x0(2, 0), np.zeros((2, 100, 1)))
    path = str(tmp_path / "ddp_trace.txt")
    ddp.dumpTraceDataList(path, instance=1)
    tab = np.genfromtxt(path, dtype=None, delimiter=" ", names=True)
    assert tab.dtype.names == ("iter", "cost", "lambda", "dlambda", "alpha", "k_rel_norm", "cost_update_actual",
                               "cost_update_expected", "cost_update_ratio", "duration_derivative", "duration_backward",
                               "duration_forward")
    assert len(tab) == ddp.n_trace()[1] and list(tab["iter"]) == list(range(len(tab)))
    np.testing.assert_allclose(tab["cost"], ddp.trace()[1, :len(tab), 1], rtol=1e-5)

    fm = gpu.FmpcSolver("oscillator", batch_capacity=1)
    fm.config().horizon_steps, fm.config().max_iter = 50, 4
    fm.solve_batch(0.0, np.array([[0.0, 1.0]]), _fmpc_var(fm, 1))
    path = str(tmp_path / "fmpc_trace.txt")
    fm.dumpTraceDataList(path)
    tab = np.genfromtxt(path, dtype=None, delimiter=" ", names=True)
    assert tab.dtype.names == ("iter", "kkt_error", "duration_coeff", "duration_backward", "duration_forward",
                               "duration_update")
    assert len(tab) == fm.n_trace()[0]
    np.testing.assert_allclose(tab["kkt_error"], fm.trace()[0, :len(tab), 1], rtol=1e-5)


def test_cartpole_swing_up_meets_the_reference_closed_loop_thresholds(gpu):
    """TestDDPCartPole end to end (TestDDPCartPole.cpp:291-358 with the parameters of tests/test/TestDDPCartPole.test):
    horizon 2 s / 0.01 s, max_iter 3, limits +-15 N, MPC every 4 ms, plant at 2 ms, 10 s of simulated time = 2500
    ticks, all on the device.  Instance 0 is the test's own start (hanging down, x = (0, pi, 0, 0)); the others start
    from perturbed states.  The assertions are the reference's: |pos - ref| < 100 throughout (:336), and at the end
    |pos - ref| < 1, |theta| < 0.1, |vel| < 1, |omega| < 0.1 (:351-354)."""
    N, B, ticks = 200, 16, 2500
    p = O.default_params("cartpole")
    rng = np.random.default_rng(1)
    x0 = np.tile([0.0, np.pi, 0.0, 0.0], (B, 1))
    x0[1:] += rng.uniform(-1, 1, (B - 1, 4)) * [0.5, 0.3, 0.2, 0.2]
    solver = gpu.DDPSolver("cartpole", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps, c.max_iter, c.with_input_constraint = N, 3, True
    solver.setInputLimitsFunc((np.array([-15.0]), np.array([15.0])))
    log = solver.run_mpc(0.0, x0, np.zeros((B, N, 1)), n_ticks=ticks, tick_dt=0.004, plant="sim", shift_inputs=False,
                         clamp_u0=True, sim_dt=0.002, n_substeps=2)
    x = log["x"]
    assert np.all(np.isfinite(x)) and np.all(np.abs(log["u"]) <= 15.0)
    assert np.all(np.abs(x[:, :, 0]) < 1e2)
    # theta is not wrapped by the problem: the upright equilibrium reached may be any multiple of 2 pi
    theta = (x[:, -1, 1] + np.pi) % (2 * np.pi) - np.pi
    final = np.stack([x[:, -1, 0], theta, x[:, -1, 2], x[:, -1, 3]], axis=1)
    ok = (np.abs(final[:, 0]) < 1.0) & (np.abs(final[:, 1]) < 1e-1) & (np.abs(final[:, 2]) < 1.0) & (np.abs(final[:, 3]) < 1e-1)
    assert ok[0], final[0]  # the reference's own scenario
    assert np.abs(x[0, -1, 1]) < 1e-1  # ... which ends at theta = 0 itself, as the reference asserts
    assert ok.mean() >= 0.9, final[~ok]


def test_fmpc_cartpole_swing_up_meets_the_reference_closed_loop_thresholds(gpu):
    """TestFmpcCartPole end to end (TestFmpcCartPole.cpp:286-384): horizon 2 s / 0.01 s, max_iter 5, inequality
    constraints |f| <= 15 N and |pos| <= 20 m inside the problem, MPC every 4 ms, plant at 2 ms with the K_0 feedback
    term, 10 s = 2500 ticks on the device, Variable.reset(0, 0, 0, 1, 1) at the start.  Same final thresholds as the
    DDP test (:377-380)."""
    N, B, ticks = 200, 8, 2500
    p = O.default_params("fmpc_cartpole")
    rng = np.random.default_rng(2)
    x0 = np.tile([0.0, np.pi, 0.0, 0.0], (B, 1))
    x0[1:] += rng.uniform(-1, 1, (B - 1, 4)) * [0.3, 0.2, 0.1, 0.1]
    solver = gpu.FmpcSolver("cartpole", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps, c.max_iter = N, 5
    log = solver.run_mpc(0.0, x0, _fmpc_var(solver, B), n_ticks=ticks, tick_dt=0.004, plant="sim", sim_dt=0.002,
                         n_substeps=2, feedback=True)
    x = log["x"]
    assert np.all(np.isfinite(x)) and np.all(np.abs(x[:, :, 0]) < 1e2)
    assert set(np.unique(log["status"])) <= {1, 5}
    theta = (x[:, -1, 1] + np.pi) % (2 * np.pi) - np.pi
    final = np.stack([x[:, -1, 0], theta, x[:, -1, 2], x[:, -1, 3]], axis=1)
    ok = (np.abs(final[:, 0]) < 1.0) & (np.abs(final[:, 1]) < 1e-1) & (np.abs(final[:, 2]) < 1.0) & (np.abs(final[:, 3]) < 1e-1)
    assert ok[0], final[0]
    assert ok.mean() >= 0.75, final[~ok]
    # the force constraint of the problem holds for the planned input of (almost) every tick once the iterate is feasible
    assert np.mean(np.abs(log["u"]) <= 15.0 + 1e-6) >= 0.99


def test_bipedal_full_run_meets_the_reference_thresholds(gpu):
    """TestDDPBipedal.TestCase1 over its whole 20 s (2000 ticks, horizon 3 s / 0.01 s) on the device: per tick
    |planned_zmp - ref_zmp| < 1e-2 (TestDDPBipedal.cpp:256), final CoM within 1e-2 of the final reference ZMP and at
    rest (:271-273).  max_iter is capped at 6 here (the problem is linear-quadratic and converges in 1-3 iterations;
    the reference leaves the default 500 and stops on its termination tests after as many)."""
    p = O.default_params("bipedal")
    N, ticks, dt = 300, 2000, p[0]
    solver = gpu.DDPSolver("bipedal", params=p, batch_capacity=1)
    c = solver.config()
    c.horizon_steps, c.max_iter = N, 6
    log = solver.run_mpc(0.0, np.zeros((1, 2)), np.zeros((1, N, 1)), n_ticks=ticks, tick_dt=dt, plant="model",
                         shift_inputs=True)
    assert np.all(log["status"] == 1), "every tick must converge within the cap"
    ref = np.array([_ref_zmp(k * dt) for k in range(ticks)])
    assert np.max(np.abs(log["u"][0, :, 0] - ref)) < 1e-2
    assert abs(log["x"][0, -1, 0] - _ref_zmp(ticks * dt)) < 1e-2 and abs(log["x"][0, -1, 1]) < 1e-2


def test_fmpc_oscillator_full_run_meets_the_reference_thresholds(gpu):
    """TestFmpcOscillator.SolveMpc over its whole 10 s (2000 ticks of 5 ms, horizon 4 s / 0.01 s, max_iter 3) on the
    device: status Succeeded or MaxIterationReached and constraints satisfied at every tick (:170, :180), final state
    within 1e-2 of the origin (:193-194)."""
    N, ticks, sim_dt = 400, 2000, 0.005
    p = O.default_params("fmpc_oscillator")
    solver = gpu.FmpcSolver("oscillator", params=p, batch_capacity=1)
    c = solver.config()
    c.horizon_steps, c.max_iter = N, 3
    log = solver.run_mpc(0.0, np.array([[0.0, 1.0]]), _fmpc_var(solver, 1), n_ticks=ticks, tick_dt=sim_dt, plant="sim",
                         sim_dt=sim_dt)
    assert set(np.unique(log["status"])) <= {1, 5}
    x, u = log["x"][0], log["u"][0, :, 0]
    g = np.stack([-x[:-1, 1] - 0.05, -u - 1.0, u - 0.9], axis=1)  # ineqConst (TestFmpcOscillator.cpp:68-76)
    assert np.all(g <= 0), g[np.any(g > 0, axis=1)][:3]
    assert abs(x[-1, 0]) < 1e-2 and abs(x[-1, 1]) < 1e-2
