"""GPU parity tests: the CUDA FMPC path (through the C ABI) and the remaining problem functors against
the CPU oracle.

FMPC tolerance: per-iteration comparison of (x, u, lambda, s, nu) after 1, 2, ..., 10
iterations, relative 1e-8; cold-started FMPC does not converge within 10 iterations, so
the iterates themselves are what is compared."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

REL_TOL = 1e-8


def _rel(a, b):
    """max |a - b| / (1 + max |b|) per instance."""
    ax = tuple(range(1, a.ndim))
    return np.max(np.abs(a - b), axis=ax) / (1.0 + np.max(np.abs(b), axis=ax))


def _initial_variable(solver, B):
    v = solver.make_variable(B)
    v.reset(0.0, 0.0, 0.0, 1.0, 1.0)  # TestFmpcCartPole.cpp:329-330
    return v


def _as_dict(v):
    return {"x": v.x_list, "u": v.u_list, "lambda": v.lambda_list, "s": v.s_list, "nu": v.nu_list}


@pytest.mark.parametrize("gpu_name,oracle_name", [("oscillator", "fmpc_oscillator"), ("cartpole", "fmpc_cartpole"),
                                                  ("bipedal", "bipedal")])
def test_functors_match_oracle(gpu, gpu_name, oracle_name):
    nx, nu, ng, _ = O.model_dims(oracle_name)
    rng = np.random.default_rng(3)
    n = 32
    x = rng.uniform(-2, 2, (n, nx))
    u = rng.uniform(-2, 2, (n, nu))
    t = rng.uniform(0, 20, n)
    p = O.default_params(oracle_name)
    d = gpu.model_eval(gpu_name, t, x, u, params=p)
    for i in range(n):
        o = O.model_eval(oracle_name, p, t[i], x[i], u[i])
        keys = ["x_next", "Fx", "Fu", "Lx", "Lu", "Lxx", "Luu", "Lxu", "Vx", "Vxx"] + (["g", "C", "D"] if ng else [])
        for key in keys:
            np.testing.assert_allclose(d[key][i], o[key], rtol=1e-12, atol=1e-13, err_msg=f"{gpu_name}.{key}")
        assert abs(d["running_cost"][i] - o["running_cost"]) <= 1e-12 * max(1.0, abs(o["running_cost"]))
        assert abs(d["terminal_cost"][i] - o["terminal_cost"]) <= 1e-12 * max(1.0, abs(o["terminal_cost"]))


@pytest.mark.parametrize("max_iter", [1, 2, 3, 5, 10])
def test_fmpc_cartpole_per_iteration_parity(gpu, max_iter):
    """BASELINE.json configs[2] shape (n_x=4, n_u=1, n_g=4: +-15 N and +-20 m, N=100) on a smaller batch."""
    B, N = 96, 100
    x0 = O.cartpole_x0(B, 3)
    p = O.default_params("fmpc_cartpole")
    solver = gpu.FmpcSolver("cartpole", params=p, batch_capacity=B)
    solver.config().max_iter = max_iter
    var = _initial_variable(solver, B)
    status = solver.solve_batch(0.0, x0, var)
    ref = O.fmpc_solve_batch("fmpc_cartpole", p, O.fmpc_config(max_iter=max_iter, horizon_steps=N), x0, _as_dict(var))
    assert np.array_equal(status, ref["status"])
    assert np.array_equal(solver.n_trace(), ref["n_trace"])
    out = _as_dict(solver.variable())
    within = np.ones(B, dtype=bool)
    for key in ("x", "u", "lambda", "s", "nu"):
        within &= _rel(out[key], ref[key]) <= REL_TOL
    assert within.mean() == 1.0, f"fraction within tolerance {within.mean():.3f}"
    tr = solver.trace()
    np.testing.assert_allclose(tr[:, :, 1], ref["trace"][:, :, 1], rtol=1e-7)  # kkt_error (derived; iterates are the 1e-8 gate)
    np.testing.assert_allclose(tr[:, :, 2:], ref["trace"][:, :, 2:], rtol=1e-7)  # barrier_eps, alpha_s, alpha_nu
    assert _rel(solver.K_list().reshape(B, N, -1), ref["K"]).max() <= 1e-7


def test_fmpc_config3_full_batch(gpu):
    """BASELINE.json configs[2] at full size: batch 1024, horizon 100, max_iter 10, seed 3."""
    B, N = 1024, 100
    x0 = O.cartpole_x0(B, 3)
    p = O.default_params("fmpc_cartpole")
    solver = gpu.FmpcSolver("cartpole", params=p, batch_capacity=B)
    var = _initial_variable(solver, B)
    status = solver.solve_batch(0.0, x0, var)
    ref = O.fmpc_solve_batch("fmpc_cartpole", p, O.fmpc_config(horizon_steps=N), x0, _as_dict(var))
    assert np.array_equal(status, ref["status"])
    out = _as_dict(solver.variable())
    err = np.max([_rel(out[k], ref[k]) for k in out], axis=0)
    frac = float((err <= REL_TOL).mean())
    print(f"FMPC config 3: fraction of instances within {REL_TOL:g}: {frac:.4f}; max rel error {err.max():.2e}")
    # "report fraction of instances within tolerance" -- measured 0.998 (2 of 1024 instances
    # amplify rounding noise to 1.3e-7 through ill-conditioned interior-point steps); nothing may be far off
    assert frac >= 0.99
    assert err.max() <= 1e-5
    assert set(np.unique(status)) <= {1, 5}


def test_fmpc_oscillator_closed_loop(gpu):
    """TestFmpcOscillator.SolveMpc (TestFmpcOscillator.cpp:137-199), first 2 s of the 10 s run: status is
    Succeeded or MaxIterationReached at every tick (:170), constraints hold (:180), and the closed loop
    tracks the oracle's closed loop."""
    N = 400
    p = O.default_params("fmpc_oscillator")
    solver = gpu.FmpcSolver("oscillator", params=p, batch_capacity=1)
    c = solver.config()
    c.horizon_steps, c.max_iter = N, 3
    var = _initial_variable(solver, 1)
    ovar = _as_dict(var)
    ocfg = O.fmpc_config(horizon_steps=N, max_iter=3)
    x = np.array([[0.0, 1.0]])
    xo = x.copy()
    t, sim_dt = 0.0, 0.005
    for _ in range(400):
        st = solver.solve(t, x[0], var)
        assert int(st) in (1, 5)
        var = solver.variable()
        u = var.u_list[0, 0]
        g = np.array([-x[0, 1] - 0.05, -u[0] - 1.0, u[0] - 0.9])
        assert np.all(g <= 0), g
        ro = O.fmpc_solve_batch("fmpc_oscillator", p, ocfg, xo, ovar, t0=t)
        ovar = {k: ro[k] for k in ("x", "u", "lambda", "s", "nu")}
        assert abs(u[0] - ro["u"][0, 0, 0]) < 1e-6
        # plant step with sim_dt (stateEq(t, x, u, sim_dt), :186)
        def plant(xx, uu):
            xd = np.array([(1.0 - xx[1] ** 2) * xx[0] - xx[1] + uu, xx[0]])
            return xx + sim_dt * xd
        x = plant(x[0], u[0])[None]
        xo = plant(xo[0], ro["u"][0, 0, 0])[None]
        t += sim_dt
    assert np.max(np.abs(x - xo)) < 1e-6


def test_fmpc_error_behaviour(gpu):
    solver = gpu.FmpcSolver("cartpole", batch_capacity=4)
    var = _initial_variable(solver, 2)
    var.s_list[1, 7, 2] = -1e-3
    with pytest.raises(gpu.NmpcB200Error) as e:  # checkVariable: std::runtime_error (FmpcSolver.hpp:351-355)
        solver.solve_batch(0.0, np.zeros((2, 4)), var)
    assert e.value.code == 2 and "must be non-negative" in e.value.message
    bad = solver.make_variable(2)
    bad.u_list = bad.u_list[:, :-1]
    with pytest.raises(ValueError) as e:  # std::invalid_argument (FmpcSolver.hpp:293-297)
        solver.solve_batch(0.0, np.zeros((2, 4)), bad)
    assert "u_list length should be 100 but 99." in str(e.value)
    # NaN in the initial state => ErrorInBackward / ErrorInForward, like the oracle
    var = _initial_variable(solver, 2)
    x0 = np.zeros((2, 4))
    x0[1, 1] = np.nan
    st = solver.solve_batch(0.0, x0, var)
    ref = O.fmpc_solve_batch("fmpc_cartpole", O.default_params("fmpc_cartpole"), O.fmpc_config(), x0, _as_dict(var))
    assert np.array_equal(st, ref["status"]) and st[1] in (2, 3, 4)


@pytest.mark.parametrize("from_multipliers", [False, True])
@pytest.mark.parametrize("max_iter", [1, 3, 6])
def test_fmpc_merit_function_line_search(gpu, max_iter, from_multipliers):
    """enable_line_search (updateVariables, FmpcSolver.hpp:755-793; setupMeritFunc / calcMeritFunc :837-982;
    l1NormDirectionalDeriv, MathUtils.h:17-38) with both constraint-scale rules ((18.33) and (18.32) of
    Nocedal & Wright): iterates, accepted alpha_s and status against the oracle."""
    B, N = 64, 100
    x0 = O.cartpole_x0(B, 13)
    p = O.default_params("fmpc_cartpole")
    solver = gpu.FmpcSolver("cartpole", params=p, batch_capacity=B)
    c = solver.config()
    c.max_iter, c.enable_line_search, c.merit_const_scale_from_lagrange_multipliers = max_iter, True, from_multipliers
    var = _initial_variable(solver, B)
    status = solver.solve_batch(0.0, x0, var)
    ocfg = O.fmpc_config(max_iter=max_iter, horizon_steps=N, enable_line_search=1,
                         merit_const_scale_from_lagrange_multipliers=int(from_multipliers))
    ref = O.fmpc_solve_batch("fmpc_cartpole", p, ocfg, x0, _as_dict(var))
    assert np.array_equal(status, ref["status"])
    tr = solver.trace()
    # the accepted step lengths are powers of 1/2 times the fraction-to-boundary value: compare them first
    same_alpha = np.all(np.abs(tr[:, :, 3] - ref["trace"][:, :, 3]) <= 1e-9 * np.maximum(1.0, ref["trace"][:, :, 3]), axis=1)
    out = _as_dict(solver.variable())
    within = np.ones(B, dtype=bool)
    for key in ("x", "u", "lambda", "s", "nu"):
        within &= _rel(out[key], ref[key]) <= REL_TOL
    # an Armijo test decided at rounding level may flip for an isolated instance; everything else must agree
    assert same_alpha.mean() >= 0.97, same_alpha.mean()
    assert within[same_alpha].all(), f"{(~within[same_alpha]).sum()} instances with equal alpha_s differ"
    # and the search really backtracks somewhere, otherwise this test checks nothing
    if max_iter >= 3:
        ftb = O.fmpc_solve_batch("fmpc_cartpole", p, O.fmpc_config(max_iter=max_iter, horizon_steps=N), x0, _as_dict(var))
        assert np.any(np.abs(ftb["trace"][:, :, 3] - ref["trace"][:, :, 3]) > 1e-6)


def test_fmpc_init_complementary_variable(gpu):
    B, N = 32, 50
    x0 = O.cartpole_x0(B, 8)
    p = O.default_params("fmpc_cartpole")
    solver = gpu.FmpcSolver("cartpole", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps, c.max_iter, c.init_complementary_variable = N, 4, True
    var = _initial_variable(solver, B)
    st = solver.solve_batch(0.0, x0, var)
    ref = O.fmpc_solve_batch("fmpc_cartpole", p,
                             O.fmpc_config(horizon_steps=N, max_iter=4, init_complementary_variable=1), x0,
                             _as_dict(var))
    assert np.array_equal(st, ref["status"])
    out = _as_dict(solver.variable())
    for k in out:
        assert _rel(out[k], ref[k]).max() <= REL_TOL, k


def test_ddp_bipedal_parity_and_receding_horizon(gpu):
    """TestDDPBipedal.TestCase1 (TestDDPBipedal.cpp:161-273): time-varying LTV problem, N=300, default
    max_iter=500, receding-horizon warm start with shifted u_list (:262-267); first 60 ticks, both
    engines in lock step, plus the test's own per-tick threshold |zmp - ref| < 1e-2 (:256)."""
    N = 300
    p = O.default_params("bipedal")
    solver = gpu.DDPSolver("bipedal", params=p, batch_capacity=1)
    solver.config().horizon_steps = N
    ocfg = O.ddp_config(horizon_steps=N)
    t, x, u = 0.0, np.zeros((1, 2)), np.zeros((1, N, 1))
    for tick in range(60):
        solver.solve_batch(t, x, u)
        ro = O.ddp_solve_batch("bipedal", p, ocfg, x, u, t0=t)
        cd = solver.controlData()
        assert solver.iterations()[0] == ro["iters"][0], tick
        assert np.max(np.abs(cd.u_list - ro["u"])) <= 1e-9 * (1 + np.max(np.abs(ro["u"])))
        assert abs(cd.u_list[0, 0, 0] - 0.0) < 1e-2  # ref_zmp(t) == 0 for t <= 1.5
        t += p[0]
        x = ro["x"][:, 1, :].copy()
        u = np.concatenate([ro["u"][:, 1:, :], ro["u"][:, -1:, :]], axis=1)


@pytest.mark.parametrize("N", [1, 2, 5])
def test_fmpc_tiny_horizons(gpu, N):
    """Horizons shorter than the loader-warp rings of F2 / F3."""
    B = 3
    x0 = O.cartpole_x0(B, 50 + N) * 0.2
    p = O.default_params("fmpc_cartpole")
    solver = gpu.FmpcSolver("cartpole", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps, c.max_iter = N, 4
    var = _initial_variable(solver, B)
    status = solver.solve_batch(0.0, x0, var)
    ref = O.fmpc_solve_batch("fmpc_cartpole", p, O.fmpc_config(max_iter=4, horizon_steps=N), x0, _as_dict(var))
    assert np.array_equal(status, ref["status"])
    out = _as_dict(solver.variable())
    for key in ("x", "u", "lambda", "s", "nu"):
        assert _rel(out[key], ref[key]).max() <= REL_TOL, key
