"""CPU tests pinning the oracle restatement of nmpc_ddp (BoxQP, models, DDPSolver)."""
import numpy as np
import pytest

import oracle_lib as O

# --- the reference's own known-answer tests: nmpc_ddp/tests/src/TestBoxQP.cpp:39-55 (tol 1e-6, :29) ---
BOXQP_KATS = [
    ([1.5, 1.0], [-10, -10], [10, 10], [-1.5, -2.0]),
    ([1.5, 1.0], [0.5, -2.0], [5.0, 2.0], [0.5, -2.0]),
    ([1.0, 1.5], [0.0, -1.0], [5.0, -0.5], [0.0, -1.0]),
    ([1.5, 1.0], [-5.0, -1.0], [-2.0, 2.0], [-2.0, -1.0]),
    ([1.0, 1.5], [-5.0, -10.0], [-2.0, 10.0], [-2.0, -3.0]),
]


@pytest.mark.parametrize("g,lower,upper,x_gt", BOXQP_KATS)
def test_boxqp_known_answers(g, lower, upper, x_gt):
    H = np.array([[1.0, 0.0], [0.0, 0.5]])
    x, retval, _ = O.boxqp_solve(H, g, lower, upper)
    assert retval > 0
    assert np.linalg.norm(x - np.array(x_gt)) < 1e-6


def test_boxqp_scalar_clamps():
    # n = 1: the constrained DDP cart-pole case degenerates to a clamped Newton step
    x, retval, _ = O.boxqp_solve(np.array([[2.0]]), [-10.0], [-1.0], [1.0])
    assert retval == 6 and x[0] == 1.0
    x, retval, _ = O.boxqp_solve(np.array([[2.0]]), [1.0], [-1.0], [1.0])
    assert retval in (4, 5) and abs(x[0] + 0.5) < 1e-12


def test_boxqp_indefinite_hessian_fails():
    _, retval, _ = O.boxqp_solve(np.array([[-1.0]]), [0.3], [-1.0], [1.0])
    assert retval == -1


def _central_diff(model, p, t, x, u, eps=1e-6):
    nx, nu = len(x), len(u)
    Fx, Fu = np.zeros((nx, nx)), np.zeros((nx, nu))
    for i in range(nx):
        d = np.zeros(nx)
        d[i] = eps
        Fx[:, i] = (O.model_eval(model, p, t, x + d, u)["x_next"] - O.model_eval(model, p, t, x - d, u)["x_next"]) / (
            2 * eps)
    for i in range(nu):
        d = np.zeros(nu)
        d[i] = eps
        Fu[:, i] = (O.model_eval(model, p, t, x, u + d)["x_next"] - O.model_eval(model, p, t, x, u - d)["x_next"]) / (
            2 * eps)
    return Fx, Fu


def test_cartpole_derivative_check():
    """TestDDPCartPole.CheckDerivative (TestDDPCartPole.cpp:609-649): x=(1,-2,3,-4), u=10, eps 1e-6, tol 1e-6."""
    p = O.default_params("cartpole")
    x, u = np.array([1.0, -2.0, 3.0, -4.0]), np.array([10.0])
    a = O.model_eval("cartpole", p, 0.0, x, u)
    Fx, Fu = _central_diff("cartpole", p, 0.0, x, u)
    assert np.linalg.norm(a["Fx"] - Fx) < 1e-6
    assert np.linalg.norm(a["Fu"] - Fu) < 1e-6


def test_fmpc_models_derivative_check():
    """TestFmpcCartPole.cpp:625-693 / TestFmpcOscillator.cpp:203-266 pattern, incl. C and D."""
    rng = np.random.default_rng(1)
    for model, x, u in (("fmpc_cartpole", np.array([1.0, -2.0, 3.0, -4.0]), np.array([10.0])),
                        ("fmpc_oscillator", rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 1))):
        p = O.default_params(model)
        a = O.model_eval(model, p, 0.0, x, u)
        Fx, Fu = _central_diff(model, p, 0.0, x, u)
        assert np.linalg.norm(a["Fx"] - Fx) < 1e-6
        assert np.linalg.norm(a["Fu"] - Fu) < 1e-6
        eps = 1e-6
        Cn = np.zeros_like(a["C"])
        for i in range(len(x)):
            d = np.zeros(len(x))
            d[i] = eps
            Cn[:, i] = (O.model_eval(model, p, 0.0, x + d, u)["g"] - O.model_eval(model, p, 0.0, x - d, u)["g"]) / (
                2 * eps)
        assert np.linalg.norm(a["C"] - Cn) < 1e-6


def test_cartpole_swingup_known_trace():
    """Regression values of SURVEY.md App. C (independent numpy probe of the reference algorithm):
    cart-pole N=100, x0=(0,pi,0,0), u_init=0, max_iter=10, reference termination."""
    p = O.default_params("cartpole")
    cfg = O.ddp_config(max_iter=10, horizon_steps=100)
    r = O.ddp_solve_batch("cartpole", p, cfg, np.array([[0, np.pi, 0, 0]]), np.zeros((1, 100, 1)))
    n = r["n_trace"][0]
    cost = r["trace"][0, :n, 1]
    want = [498.415022255, 406.088004263, 404.897899816, 404.865142163, 404.86399135, 404.863948982, 404.863947391,
            404.863947331]
    assert n == 8 and r["status"][0] == 1 and r["iters"][0] == 7
    np.testing.assert_allclose(cost, want, rtol=2e-12)
    np.testing.assert_allclose(r["trace"][0, :n, 2],
                               [1e-4, 6.25e-5, 2.44140625e-5, 5.9604644775390625e-6, 9.094947017729282e-07, 0, 0, 0],
                               rtol=1e-12)
    assert np.all(r["trace"][0, 1:n, 4] == 1.0)
    np.testing.assert_allclose(r["u"][0, :4, 0],
                               [18.786169762863928, 18.484618096972063, 18.184692098150514, 17.886397315575014],
                               rtol=1e-10)
    assert r["n_fwd"][0] == 7 and r["n_bwd"][0] == 7


def test_cartpole_batch_iteration_histogram():
    """200 random x0 (seed 0): 197 converge within 10 iterations, histogram of iterations."""
    p = O.default_params("cartpole")
    cfg = O.ddp_config(max_iter=10, horizon_steps=100)
    r = O.ddp_solve_batch("cartpole", p, cfg, O.cartpole_x0(200, 0), np.zeros((200, 100, 1)))
    hist = np.bincount(r["iters"], minlength=11)
    assert list(hist[3:]) == [2, 22, 25, 33, 73, 38, 3, 4]
    assert int((r["status"] == 1).sum()) == 197


def test_cartpole_constrained_config1():
    """BASELINE.json configs[0]: TestDDPCartPole first MPC tick (N=200, max_iter=3, +-15 N, BoxQP branch);
    SURVEY.md App. C probe values.  N=400 (the literal config text) must run as well."""
    p = O.default_params("cartpole")
    lo, hi = np.array([-15.0]), np.array([15.0])
    cfg = O.ddp_config(max_iter=3, horizon_steps=200, with_input_constraint=1)
    r = O.ddp_solve_batch("cartpole", p, cfg, np.array([[0, np.pi, 0, 0]]), np.zeros((1, 200, 1)), u_lo=lo, u_hi=hi)
    np.testing.assert_allclose(r["trace"][0, :, 1], [991.8952423095, 882.7146741324, 850.4060100400, 843.1435488411],
                               rtol=1e-11)
    assert abs(r["u"][0, 0, 0] - (-8.279420)) < 1e-6
    assert np.all(r["u"][0] <= 15.0 + 1e-9) and np.all(r["u"][0] >= -15.0 - 1e-9)
    cfg = O.ddp_config(max_iter=3, horizon_steps=400, with_input_constraint=1)
    r = O.ddp_solve_batch("cartpole", p, cfg, np.array([[0, np.pi, 0, 0]]), np.zeros((1, 400, 1)), u_lo=lo, u_hi=hi)
    cost = r["trace"][0, :, 1]
    assert np.all(np.diff(cost) < 0) and r["status"][0] == 0


def test_initial_u_list_length_is_checked():
    p = O.default_params("cartpole")
    cfg = O.ddp_config(max_iter=1, horizon_steps=50)
    with pytest.raises(Exception):
        # 40 steps of input for a 50-step horizon: the wrapper reshapes, the size check fires
        O.ddp_solve_batch("cartpole", p, cfg, np.zeros((1, 4)), np.zeros((1, 40, 1)))


def test_bipedal_closed_loop_thresholds():
    """TestDDPBipedal.TestCase1 (TestDDPBipedal.cpp:161-273), shortened to the first 3 s of the 20 s run:
    per-tick |planned_zmp - ref_zmp| < 1e-2 with receding-horizon warm start."""
    p = O.default_params("bipedal")
    N = 300
    cfg = O.ddp_config(horizon_steps=N)
    dt = p[0]
    t, x, u = 0.0, np.zeros((1, 2)), np.zeros((1, N, 1))

    def ref_zmp(tt):
        tt += 1e-6
        if tt <= 1.5 or tt >= 20.0 - 1.5:
            return 0.0
        return 0.15 if int(np.floor((tt - 1.0) / 1.0)) % 2 == 0 else -0.15

    for _ in range(300):
        r = O.ddp_solve_batch("bipedal", p, cfg, x, u, t0=t)
        assert abs(r["u"][0, 0, 0] - ref_zmp(t)) < 1e-2
        t += dt
        x = r["x"][:, 1, :].copy()
        u = np.concatenate([r["u"][:, 1:, :], r["u"][:, -1:, :]], axis=1)
