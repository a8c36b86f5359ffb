"""Time-varying input dimension (DDPProblem<StateDim, Eigen::Dynamic>, DDPProblem.h:61-85).

Golden vectors: tests/golden/reference_vertical.npz, produced by tests/golden/make_golden_vertical.py from the
REFERENCE's own DDPSolver<2, Eigen::Dynamic> (and BoxQP<Dynamic>) on the problem and MPC loop of
nmpc_ddp/tests/src/TestDDPVerticalMotion.cpp.  The oracle and the device keep compile-time sizes (NU = the largest
dimension) and treat inputs a >= inputDim(t) as decoupled padding; these tests pin that construction against the
reference's reduced-dimension results: CPU oracle here, CUDA path under -m gpu."""
import os

import numpy as np
import pytest

import oracle_lib as O
import ref_lib as R

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vertical.npz"))
N = int(GOLDEN["N"])
DT = 0.01
LO, HI = np.zeros(2), np.full(2, 30.0)


def input_dim(t):
    t += 1e-6  # TestDDPVerticalMotion.cpp:61-78
    if 2.0 < t < 3.0:
        return 2
    if 4.5 < t < 5.0:
        return 0
    return 1


def shifted_warm_start(u, t):
    """TestDDPVerticalMotion.cpp:303-315 on padded arrays: drop u_list[0]; the new last entry repeats the old one when
    the dimension at the new terminal time is the same, else it is Zero(terminal_input_dim)."""
    last = u[:, -1:].copy()
    if input_dim(t + (N - 1) * DT) != input_dim(t + N * DT):
        last[...] = 0.0
    return np.concatenate([u[:, 1:], last], axis=1)


def oracle_loop(with_constraint, ticks):
    p = O.default_params("vertical_motion")
    x, u, t = np.array([[1.2, 0.0]]), np.zeros((1, N, 2)), 0.0
    log = {"x": [], "u0": [], "iters": []}
    kw = dict(horizon_steps=N, initial_lambda=1e-6, with_input_constraint=int(with_constraint))
    for k in range(ticks):
        cfg = O.ddp_config(max_iter=500 if k == 0 else 3, **kw)  # :287
        r = O.ddp_solve_batch("vertical_motion", p, cfg, x, u, t0=t, u_lo=LO, u_hi=HI)
        log["x"].append(x[0].copy())
        log["u0"].append(r["u"][0, 0].copy())
        log["iters"].append(int(r["iters"][0]))
        u = shifted_warm_start(r["u"], t)
        x = r["x"][:, 1].copy()
        t = (k + 1) * DT
    return {k: np.array(v) for k, v in log.items()}, r


def test_padding_pattern_of_the_golden_vectors():
    """Sanity of the fixture itself: the reference's u_list sizes follow inputDim(t) and padded entries are zero."""
    for tag in ("free", "box"):
        u = GOLDEN[f"first_{tag}/u"]
        dims = np.array([input_dim(i * DT) for i in range(N)])
        assert set(dims) == {1, 2}  # the first horizon [0, 3) s sees one and two contacts
        assert np.all(u[dims == 1, 1] == 0.0) and np.any(u[dims == 2, 1] != 0.0)
        assert np.array_equal(GOLDEN[f"loop_{tag}/dim_log"], [input_dim(k * DT) for k in range(int(GOLDEN["ticks"]))])


@pytest.mark.parametrize("tag", ["free", "box"])
def test_oracle_first_solve_matches_reference_dynamic_solver(tag):
    log, r = oracle_loop(tag == "box", 1)
    np.testing.assert_allclose(r["u"][0], GOLDEN[f"first_{tag}/u"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(r["x"][0], GOLDEN[f"first_{tag}/x"], rtol=0, atol=1e-10)
    assert log["iters"][0] == GOLDEN[f"first_{tag}/iters_log"][0]
    dims = np.array([input_dim(i * DT) for i in range(N)])
    assert np.all(r["u"][0][dims == 1, 1] == 0.0)  # the padding stays exactly zero
    if tag == "box":
        assert r["u"][0].min() >= 0.0 and r["u"][0].max() <= 30.0


@pytest.mark.parametrize("tag", ["free", "box"])
def test_oracle_mpc_loop_matches_reference_dynamic_solver(tag):
    """230 ticks: the horizon's terminal time crosses 4.5 s and 5.0 s, so the input dimension of the appended entry
    changes 1 -> 0 -> 1 and steps WITHOUT inputs enter the backward pass (DDPSolver.hpp:513-517)."""
    ticks = int(GOLDEN["ticks"])
    log, r = oracle_loop(tag == "box", ticks)
    np.testing.assert_allclose(log["x"], GOLDEN[f"loop_{tag}/x_log"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(log["u0"], GOLDEN[f"loop_{tag}/u0_log"], rtol=0, atol=1e-6)
    assert np.array_equal(log["iters"], GOLDEN[f"loop_{tag}/iters_log"])
    np.testing.assert_allclose(r["u"][0], GOLDEN[f"loop_{tag}/u"], rtol=0, atol=1e-6)
    # the reference's own check (:291-293)
    assert np.all(np.abs(log["x"][:, 0] - 1.0) < 1.0)


@pytest.mark.skipif(not R.available(), reason="needs the reference checkout (/root/reference)")
def test_reference_build_reproduces_the_golden_vectors():
    out = R.vertical_mpc(N, 1, 12)
    np.testing.assert_array_equal(out["x_log"], GOLDEN["loop_box/x_log"][:12])
    np.testing.assert_array_equal(out["u0_log"], GOLDEN["loop_box/u0_log"][:12])


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["free", "box"])
def test_device_first_solve(gpu, tag):
    p = O.default_params("vertical_motion")
    solver = gpu.DDPSolver("vertical_motion", params=p, batch_capacity=1)
    c = solver.config()
    c.horizon_steps, c.initial_lambda, c.with_input_constraint = N, 1e-6, tag == "box"
    solver.setInputLimitsFunc((LO, HI))
    u_init = np.full((1, N, 2), 0.0)
    u_init[0, :, 1] = 123.0  # garbage in the padding of initial_u_list must not matter where it IS padding
    dims = np.array([input_dim(i * DT) for i in range(N)])
    u_init[0, dims == 2, 1] = 0.0
    solver.solve_batch(0.0, np.array([[1.2, 0.0]]), u_init)
    cd = solver.controlData()
    np.testing.assert_allclose(cd.u_list[0], GOLDEN[f"first_{tag}/u"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(cd.x_list[0], GOLDEN[f"first_{tag}/x"], rtol=0, atol=1e-10)
    assert solver.iterations()[0] == GOLDEN[f"first_{tag}/iters_log"][0]
    assert np.all(cd.u_list[0][dims == 1, 1] == 0.0)
    K = solver.K_list()[0]  # [N, NU, NX]
    assert np.all(K[dims == 1, 1, :] == 0.0) and np.all(solver.k_list()[0][dims == 1, 1] == 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["free", "box"])
def test_device_mpc_loop(gpu, tag):
    """The test's loop on the device: first solve with max_iter 500 from the host (the test lowers max_iter to 3
    after it, :287), the remaining 229 ticks with run_mpc; the dimension-aware warm-start rule runs in the kernel."""
    ticks = int(GOLDEN["ticks"])
    p = O.default_params("vertical_motion")
    B = 3
    solver = gpu.DDPSolver("vertical_motion", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps, c.initial_lambda, c.with_input_constraint = N, 1e-6, tag == "box"
    solver.setInputLimitsFunc((LO, HI))
    x0 = np.array([[1.2, 0.0], [1.0, 0.1], [0.8, -0.1]])  # instance 0 is the test's own start
    solver.solve_batch(0.0, x0, np.zeros((B, N, 2)))
    cd = solver.controlData()
    assert solver.iterations()[0] == GOLDEN[f"loop_{tag}/iters_log"][0]
    np.testing.assert_allclose(cd.u_list[0, 0], GOLDEN[f"loop_{tag}/u0_log"][0], rtol=0, atol=1e-9)
    c.max_iter = 3
    got = solver.run_mpc(DT, cd.x_list[:, 1].copy(), shifted_warm_start(cd.u_list, 0.0), n_ticks=ticks - 1, tick_dt=DT,
                         plant="model", shift_inputs=True)
    np.testing.assert_allclose(got["x"][0, :-1], GOLDEN[f"loop_{tag}/x_log"][1:], rtol=0, atol=1e-8)
    np.testing.assert_allclose(got["u"][0], GOLDEN[f"loop_{tag}/u0_log"][1:], rtol=0, atol=1e-6)
    assert np.array_equal(got["iters"][0], GOLDEN[f"loop_{tag}/iters_log"][1:])
    np.testing.assert_allclose(solver.controlData().u_list[0], GOLDEN[f"loop_{tag}/u"], rtol=0, atol=1e-6)
    # every instance keeps its padding at zero and tracks the reference height (:291-293)
    dims = np.array([input_dim((k + 1) * DT) for k in range(ticks - 1)])
    assert np.all(got["u"][:, dims < 2, 1] == 0.0) and np.all(got["u"][:, dims == 0, 0] == 0.0)
    assert np.all(np.abs(got["x"][:, :, 0] - 1.0) < 1.0)
