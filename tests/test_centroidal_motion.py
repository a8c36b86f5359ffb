"""Centroidal motion: n_x = 9, input dimension 16 or 0 along the horizon (DDPProblem<9, Eigen::Dynamic>; the
reference's TestDDPCentroidalMotion.cpp).

Golden vectors: tests/golden/reference_centroidal.npz, produced by tests/golden/make_golden_centroidal.py from the
REFERENCE's own DDPSolver<9, Eigen::Dynamic> on the problem and the 100-tick MPC loop of the test.  The oracle and the
device keep compile-time sizes (NU = 16) and treat the inputs of the flight phase as decoupled padding; these tests pin
that construction for the largest problem of the reference's test suite: CPU oracle here, CUDA path under -m gpu."""
import os

import numpy as np
import pytest

import oracle_lib as O
import ref_lib as R

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_centroidal.npz"))
N = int(GOLDEN["N"])
TICKS = int(GOLDEN["ticks"])
DT = 0.03
X0 = np.array([[0.0, 0.0, 1.0, 0, 0, 0, 0, 0, 0]])


def input_dim(t):
    t += 1e-6  # TestDDPCentroidalMotion.cpp:246-266
    return 0 if 1.4 <= t < 1.6 else 16


def ref_pos(t):
    t += 1e-6  # :267-279
    return np.array([0.0, 0.0, 1.0]) if t < 1.5 else np.array([0.5, 0.0, 1.0])


def shifted_warm_start(u, t):
    """TestDDPCentroidalMotion.cpp:325-337 on padded arrays: drop u_list[0]; the new last entry repeats the old one
    when the dimension at the new terminal time is the same, else it is Zero(terminal_input_dim)."""
    last = u[:, -1:].copy()
    if input_dim(t + (N - 1) * DT) != input_dim(t + N * DT):
        last[...] = 0.0
    return np.concatenate([u[:, 1:], last], axis=1)


def oracle_loop(ticks):
    p = O.default_params("centroidal_motion")
    x, u, t = X0.copy(), np.zeros((1, N, 16)), 0.0
    log = {"x": [], "u0": [], "iters": []}
    first = None
    for k in range(ticks):
        cfg = O.ddp_config(max_iter=500 if k == 0 else 3, horizon_steps=N)  # :303
        r = O.ddp_solve_batch("centroidal_motion", p, cfg, x, u, t0=t)
        first = first or r
        log["x"].append(x[0].copy())
        log["u0"].append(r["u"][0, 0].copy())
        log["iters"].append(int(r["iters"][0]))
        u = shifted_warm_start(r["u"], t)
        x = r["x"][:, 1].copy()
        t = t + DT  # current_t += dt (:338)
    return {k: np.array(v) for k, v in log.items()}, first, r, x


def test_fixture_follows_the_stance_schedule():
    dims = np.array([input_dim(i * DT) for i in range(N)])
    assert set(dims) == {0, 16}
    assert np.all(GOLDEN["u_first"][dims == 0] == 0.0) and np.all(np.any(GOLDEN["u_first"][dims == 16] != 0.0, axis=1))
    t, expect = 0.0, []
    for _ in range(TICKS):
        expect.append(input_dim(t))
        t += DT
    assert np.array_equal(GOLDEN["dim_log"], expect)


def test_oracle_first_solve_matches_reference_dynamic_solver():
    _, first, _, _ = oracle_loop(1)
    np.testing.assert_allclose(first["u"][0], GOLDEN["u_first"], rtol=0, atol=1e-9 * np.abs(GOLDEN["u_first"]).max())
    np.testing.assert_allclose(first["x"][0], GOLDEN["x_first"], rtol=0, atol=1e-10)
    assert first["iters"][0] == GOLDEN["iters_log"][0]
    assert abs(first["cost"][0] - GOLDEN["cost_first"][0]) <= 1e-12 * abs(GOLDEN["cost_first"][0])
    dims = np.array([input_dim(i * DT) for i in range(N)])
    assert np.all(first["u"][0][dims == 0] == 0.0)  # the padding stays exactly zero


def test_oracle_mpc_loop_matches_reference_dynamic_solver():
    """The whole test: 100 ticks; the flight phase moves through the horizon, reaches its start (16 -> 0 -> 16 at
    u_list[0]) and the appended terminal entry changes dimension twice."""
    log, _, r, x_end = oracle_loop(TICKS)
    umax = np.abs(GOLDEN["u0_log"]).max()
    np.testing.assert_allclose(log["x"], GOLDEN["x_log"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(log["u0"], GOLDEN["u0_log"], rtol=0, atol=1e-8 * umax)
    assert np.array_equal(log["iters"], GOLDEN["iters_log"])
    np.testing.assert_allclose(r["u"][0], GOLDEN["u"], rtol=0, atol=1e-8 * umax)
    # the reference's own checks (:308-311, :341-343)
    t = 0.0
    for k in range(TICKS):
        assert np.linalg.norm(log["x"][k, :3] - ref_pos(t)) < 1.0
        t += DT
    assert np.linalg.norm(x_end[0, :3] - ref_pos(t)) < 1e-2 and np.linalg.norm(x_end[0, 3:]) < 1.0


def test_oracle_derivatives_match_finite_differences():
    """TEST(TestDDPCentroidalMotion, CheckDerivative) (:355-411): analytical Fx, Fu against central differences."""
    rng = np.random.default_rng(0)
    p = O.default_params("centroidal_motion")
    x, u = rng.uniform(-1, 1, 9), rng.uniform(-1, 1, 16)
    ev = O.model_eval("centroidal_motion", p, 0.0, x, u)
    eps = 1e-6
    Fx = np.zeros((9, 9))
    Fu = np.zeros((9, 16))
    for i in range(9):
        d = np.zeros(9)
        d[i] = eps
        Fx[:, i] = (O.model_eval("centroidal_motion", p, 0.0, x + d, u)["x_next"]
                    - O.model_eval("centroidal_motion", p, 0.0, x - d, u)["x_next"]) / (2 * eps)
    for i in range(16):
        d = np.zeros(16)
        d[i] = eps
        Fu[:, i] = (O.model_eval("centroidal_motion", p, 0.0, x, u + d)["x_next"]
                    - O.model_eval("centroidal_motion", p, 0.0, x, u - d)["x_next"]) / (2 * eps)
    assert np.linalg.norm(ev["Fx"] - Fx) < 1e-6 and np.linalg.norm(ev["Fu"] - Fu) < 1e-6


@pytest.mark.skipif(not R.available(), reason="needs the reference checkout (/root/reference)")
def test_reference_build_reproduces_the_golden_vectors():
    out = R.centroidal_mpc(N, 6)
    np.testing.assert_array_equal(out["x_log"], GOLDEN["x_log"][:6])
    np.testing.assert_array_equal(out["u0_log"], GOLDEN["u0_log"][:6])
    np.testing.assert_array_equal(out["u_first"], GOLDEN["u_first"])


# ------------------------------------------------------------------------------------------- GPU
def _record(name, **values):
    """Append the measured deviations to gpurun_out/centroidal_parity.json (evidence for DESIGN.md) before asserting."""
    import json
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "centroidal_parity.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[name] = {k: (v.tolist() if isinstance(v, np.ndarray) else float(v)) for k, v in values.items()}
        json.dump(data, open(path, "w"), indent=1)
    except OSError:
        pass


@pytest.mark.gpu
def test_device_model_matches_oracle(gpu):
    """The functor evaluated on the device against the same functor on the host (all nine outputs of a step)."""
    rng = np.random.default_rng(1)
    p = O.default_params("centroidal_motion")
    for t in (0.0, 1.5, 2.0):
        x, u = rng.uniform(-1, 1, 9), rng.uniform(-50, 50, 16)
        if input_dim(t) == 0:
            u[:] = 0.0
        dev = gpu.model_eval("centroidal_motion", np.array([t]), x[None], u[None], params=p)
        ref = O.model_eval("centroidal_motion", p, t, x, u)
        for k in ("x_next", "Fx", "Fu", "Lx", "Lu", "Lxx", "Luu", "Lxu", "Vx", "Vxx", "running_cost", "terminal_cost"):
            np.testing.assert_allclose(np.asarray(dev[k])[0], ref[k], rtol=1e-13, atol=1e-13, err_msg=f"{k} at t={t}")


@pytest.mark.gpu
def test_device_first_solve(gpu):
    p = O.default_params("centroidal_motion")
    B = 5
    solver = gpu.DDPSolver("centroidal_motion", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps = N
    x0 = np.repeat(X0, B, axis=0)
    x0[1:, :3] += np.random.default_rng(2).uniform(-0.05, 0.05, (B - 1, 3))  # instance 0 is the test's own start
    u_init = np.zeros((B, N, 16))
    dims = np.array([input_dim(i * DT) for i in range(N)])
    u_init[:, dims == 0, :] = 123.0  # garbage in the padding of initial_u_list must not matter
    solver.solve_batch(0.0, x0, u_init)
    cd = solver.controlData()
    umax = np.abs(GOLDEN["u_first"]).max()
    ref = O.ddp_solve_batch("centroidal_motion", p, O.ddp_config(max_iter=500, horizon_steps=N), x0, np.zeros((B, N, 16)))
    _record("first_solve", iters_device=solver.iterations(), iters_reference=GOLDEN["iters_log"][:1], iters_oracle=ref["iters"],
            du_vs_reference_rel=np.abs(cd.u_list[0] - GOLDEN["u_first"]).max() / umax,
            dx_vs_reference=np.abs(cd.x_list[0] - GOLDEN["x_first"]).max(),
            dcost_vs_reference_rel=abs(solver.cost()[0] - GOLDEN["cost_first"][0]) / abs(GOLDEN["cost_first"][0]),
            du_vs_oracle_rel=np.abs(cd.u_list - ref["u"]).max() / umax,
            dcost_vs_oracle_rel=np.abs(solver.cost() / ref["cost"] - 1).max())
    # measured on B200: 1e-15 relative against the reference's own solver (gpurun_out/centroidal_parity.json)
    np.testing.assert_allclose(cd.u_list[0], GOLDEN["u_first"], rtol=0, atol=1e-9 * umax)
    np.testing.assert_allclose(cd.x_list[0], GOLDEN["x_first"], rtol=0, atol=1e-10)
    assert solver.iterations()[0] == GOLDEN["iters_log"][0]
    assert abs(solver.cost()[0] - GOLDEN["cost_first"][0]) <= 1e-12 * abs(GOLDEN["cost_first"][0])
    assert np.all(cd.u_list[:, dims == 0] == 0.0)
    assert np.all(solver.K_list()[:, dims == 0] == 0.0) and np.all(solver.k_list()[:, dims == 0] == 0.0)
    # the other instances against the oracle
    assert np.array_equal(solver.iterations(), ref["iters"])
    np.testing.assert_allclose(cd.u_list, ref["u"], rtol=0, atol=1e-9 * umax)
    np.testing.assert_allclose(solver.cost(), ref["cost"], rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_device_mpc_loop(gpu):
    """The test's loop on the device: first solve with max_iter 500 from the host (the test lowers max_iter to 3 after
    it, :303), the remaining 99 ticks with run_mpc; the dimension-aware warm-start rule runs in the kernel.  Ends with
    the reference's own acceptance thresholds (:341-343)."""
    p = O.default_params("centroidal_motion")
    B = 2
    solver = gpu.DDPSolver("centroidal_motion", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps = N
    x0 = np.repeat(X0, B, axis=0)
    x0[1, 0] += 0.02
    solver.solve_batch(0.0, x0, np.zeros((B, N, 16)))
    cd = solver.controlData()
    assert solver.iterations()[0] == GOLDEN["iters_log"][0]
    c.max_iter = 3
    got = solver.run_mpc(DT, cd.x_list[:, 1].copy(), shifted_warm_start(cd.u_list, 0.0), n_ticks=TICKS - 1, tick_dt=DT,
                         plant="model", shift_inputs=True)
    umax = np.abs(GOLDEN["u0_log"]).max()
    _record("mpc_loop", iters_equal=float(np.array_equal(got["iters"][0], GOLDEN["iters_log"][1:])),
            dx_vs_reference=np.abs(got["x"][0, :-1] - GOLDEN["x_log"][1:]).max(),
            du0_vs_reference_rel=np.abs(got["u"][0] - GOLDEN["u0_log"][1:]).max() / umax,
            du_last_vs_reference_rel=np.abs(solver.controlData().u_list[0] - GOLDEN["u"]).max() / umax,
            x_end=got["x"][:, -1])
    np.testing.assert_allclose(got["x"][0, :-1], GOLDEN["x_log"][1:], rtol=0, atol=1e-9)
    np.testing.assert_allclose(got["u"][0], GOLDEN["u0_log"][1:], rtol=0, atol=1e-8 * umax)
    assert np.array_equal(got["iters"][0], GOLDEN["iters_log"][1:])
    np.testing.assert_allclose(solver.controlData().u_list[0], GOLDEN["u"], rtol=0, atol=1e-8 * umax)
    dims = np.array([input_dim((k + 1) * DT) for k in range(TICKS - 1)])
    assert np.all(got["u"][:, dims == 0] == 0.0)
    x_end = got["x"][:, -1]
    assert np.all(np.linalg.norm(x_end[:, :3] - ref_pos(TICKS * DT), axis=1) < 1e-2 + 0.03)  # instance 1 starts 2 cm off
    assert np.linalg.norm(x_end[0, :3] - ref_pos(TICKS * DT)) < 1e-2 and np.all(np.linalg.norm(x_end[:, 3:], axis=1) < 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("reg_type", [1, 2])
def test_device_wide_sweep_agrees_with_the_cooperative_sweep_and_the_oracle(gpu, reg_type, monkeypatch):
    """K2 for many inputs (ddp_backward_wide.cuh, the default at n_u = 16) against the cooperative variant
    (knob backward_wide = 0, here through its environment preset NMPC_B200_BWD_WIDE) and the oracle, both regularisation types, a ragged batch (odd number of instances: the
    second half of the last warp idles)."""
    p = O.default_params("centroidal_motion")
    B = 7
    x0 = np.repeat(X0, B, axis=0)
    x0[:, :3] += np.random.default_rng(5).uniform(-0.05, 0.05, (B, 3))
    u_init = np.zeros((B, N, 16))
    kw = dict(max_iter=6, horizon_steps=N, reg_type=reg_type)
    ref = O.ddp_solve_batch("centroidal_motion", p, O.ddp_config(**kw), x0, u_init)
    out = {}
    for tag, env in (("wide", None), ("coop", "0")):
        if env is not None:
            monkeypatch.setenv("NMPC_B200_BWD_WIDE", env)
        solver = gpu.DDPSolver("centroidal_motion", params=p, batch_capacity=B)
        c = solver.config()
        c.horizon_steps, c.max_iter, c.reg_type = N, 6, reg_type
        solver.solve_batch(0.0, x0, u_init)
        out[tag] = (solver.controlData().u_list.copy(), solver.cost().copy(), solver.iterations().copy(),
                    solver.K_list().copy(), solver.n_backward().copy())
        solver.close()
    umax = np.abs(ref["u"]).max()
    _record(f"wide_vs_coop_reg{reg_type}", du_wide_vs_oracle_rel=np.abs(out["wide"][0] - ref["u"]).max() / umax,
            du_coop_vs_oracle_rel=np.abs(out["coop"][0] - ref["u"]).max() / umax,
            du_wide_vs_coop_rel=np.abs(out["wide"][0] - out["coop"][0]).max() / umax,
            dK_wide_vs_coop=np.abs(out["wide"][3] - out["coop"][3]).max(), Kmax=np.abs(out["coop"][3]).max(),
            dcost_wide_vs_oracle_rel=np.abs(out["wide"][1] / ref["cost"] - 1).max())
    for tag in ("wide", "coop"):
        np.testing.assert_allclose(out[tag][0], ref["u"], rtol=0, atol=1e-9 * umax, err_msg=tag)
        # six iterations are mid-convergence: with reg_type 2 the cost of this ill-conditioned problem (input weight
        # 1e-6) carries the FMA-contraction differences at 2e-11 (measured; 4e-14 with reg_type 1)
        np.testing.assert_allclose(out[tag][1], ref["cost"], rtol=1e-12 if reg_type == 1 else 1e-9, atol=0, err_msg=tag)
        assert np.array_equal(out[tag][2], ref["iters"]) and np.array_equal(out[tag][4], ref["n_bwd"])
    # same expressions in the same order: the two device variants agree to the last bit
    assert np.array_equal(out["wide"][0], out["coop"][0]) and np.array_equal(out["wide"][3], out["coop"][3])


@pytest.mark.gpu
def test_device_with_input_limits_many_inputs(gpu):
    """n_u = 16 WITH input limits (with_input_constraint: the cooperative K2 with the in-register BoxQP of 16 variables;
    the reference's centroidal test is unconstrained, TestDDPCentroidalMotion.cpp:247).  Ridge forces in [0, f_max]:
    the lower bound is active wherever the unconstrained solution would pull.  Against the oracle (whose BoxQP passes
    the reference's known-answer tests): same iteration / backward-pass counts and status, trajectories and costs at
    the M-ref tolerances over the first two iterations -- from then on an input sitting ON a bound makes the
    clamped-set test an exact floating-point equality (BoxQP.h:189-191) -- and the bounds respected throughout."""
    p = O.default_params("centroidal_motion")
    B = 4
    x0 = np.repeat(X0, B, axis=0)
    x0[1:, :3] += np.random.default_rng(11).uniform(-0.05, 0.05, (B - 1, 3))
    u_init = np.zeros((B, N, 16))
    lo, hi = np.zeros(16), np.full(16, 40.0)
    for max_iter in (1, 2, 6):
        ref = O.ddp_solve_batch("centroidal_motion", p, O.ddp_config(max_iter=max_iter, horizon_steps=N, with_input_constraint=1),
                                x0, u_init, u_lo=lo, u_hi=hi)
        solver = gpu.DDPSolver("centroidal_motion", params=p, batch_capacity=B)
        c = solver.config()
        c.horizon_steps, c.max_iter, c.with_input_constraint = N, max_iter, True
        solver.setInputLimitsFunc((lo, hi))
        solver.solve_batch(0.0, x0, u_init)
        u = solver.controlData().u_list
        umax = np.abs(ref["u"]).max()
        _record(f"boxqp_nu16_it{max_iter}", du_vs_oracle_rel=np.abs(u - ref["u"]).max() / umax,
                dcost_vs_oracle_rel=np.abs(solver.cost() / ref["cost"] - 1).max(), iters_device=solver.iterations(),
                iters_oracle=ref["iters"], clamped_fraction=float(np.mean(solver.K_list().reshape(B, N, 16, 9).any(axis=3) == 0)))
        assert np.array_equal(solver.status(), ref["status"])
        if max_iter <= 2:
            assert np.array_equal(solver.iterations(), ref["iters"]) and np.array_equal(solver.n_backward(), ref["n_bwd"])
            np.testing.assert_allclose(u, ref["u"], rtol=0, atol=1e-9 * umax)
            np.testing.assert_allclose(solver.cost(), ref["cost"], rtol=1e-10, atol=0)
        else:
            np.testing.assert_allclose(solver.cost(), ref["cost"], rtol=1e-2, atol=0)
        dims = np.array([input_dim(i * DT) for i in range(N)])
        assert np.all(u[:, dims == 0] == 0.0)
        solver.close()
