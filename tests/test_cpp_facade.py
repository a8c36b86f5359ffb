"""The C++ template facades (include/nmpc_ddp, include/nmpc_fmpc): compiled with g++ against the C ABI
and driven the way the reference's tests drive DDPSolver / FmpcSolver."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_facade")


@pytest.fixture(scope="module")
def facade_bin(nmpc):
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_facade.cpp"), "-o", BIN, "-L" + os.path.join(ROOT, "nmpc_b200"),
           "-lnmpc_b200", "-Wl,-rpath," + os.path.join(ROOT, "nmpc_b200")]
    subprocess.run(cmd, check=True)
    return BIN


def _parse(out):
    d = {}
    for line in out.splitlines():
        k, _, v = line.partition(" ")
        d[k] = v
    return d


def test_facade_compiles_and_fails_loudly_without_gpu(facade_bin, nmpc):
    """Host-only part: derivative check through the DDPProblem virtuals (TestDDPCartPole.cpp:609-649)."""
    r = subprocess.run([facade_bin, "--host-only"], capture_output=True, text=True)
    d = _parse(r.stdout)
    ex, eu = (float(v) for v in d["deriv_err"].split())
    assert ex < 1e-6 and eu < 1e-6 and r.returncode == 0
    # TestDDPCentroidalMotion.CheckDerivative (:355-411) and inputDim(t) of the Dynamic-dimension problems
    ex, eu = (float(v) for v in d["centroidal_deriv_err"].split())
    assert ex < 1e-6 and eu < 1e-6
    assert d["input_dims"].split() == ["16", "0", "16", "1", "2", "0"]
    if nmpc.device_count() == 0:
        assert "no CPU fallback" in d["host_only_error"]


@pytest.mark.gpu
def test_facade_against_oracle(facade_bin, gpu):
    r = subprocess.run([facade_bin], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    d = _parse(r.stdout)
    p = O.default_params("cartpole")
    x0 = np.array([[0, np.pi, 0, 0]])
    ref = O.ddp_solve_batch("cartpole", p, O.ddp_config(max_iter=10), x0, np.zeros((1, 100, 1)))
    assert d["ddp_converged"] == "1"
    cost = np.array([float(v) for v in d["ddp_trace_cost"].split()])
    np.testing.assert_allclose(cost, ref["trace"][0, :ref["n_trace"][0], 1], rtol=1e-12)
    u = np.array([float(v) for v in d["ddp_u"].split()])
    assert np.max(np.abs(u - ref["u"][0, :, 0])) <= 1e-9 * (1 + np.max(np.abs(u)))
    assert abs(float(d["ddp_cost_sum"]) - ref["cost"][0]) <= 1e-12 * ref["cost"][0]
    assert int(d["ddp_warm_iters"].split()[0]) <= 2
    assert d["ddp_invalid_argument"] == "initial_u_list length should be 100 but 99."
    durations = [float(v) for v in d["ddp_duration_ms"].split()]
    assert all(v > 0 for v in durations)
    header = open("/tmp/nmpc_b200_TestDDPCartPoleTraceData.txt").readline().split()
    assert header == ["iter", "cost", "lambda", "dlambda", "alpha", "k_rel_norm", "cost_update_actual",
                      "cost_update_expected", "cost_update_ratio", "duration_derivative", "duration_backward",
                      "duration_forward"]  # the columns scripts/plotDDPTraceData.py reads

    # runMpc: 5 ticks of TestDDPCartPole's loop for 2 instances vs the same loop driven from here around the oracle
    lo, hi = np.array([-15.0]), np.array([15.0])
    cfg = O.ddp_config(horizon_steps=200, max_iter=3, with_input_constraint=1)
    p_sim = p.copy()
    p_sim[0] = 0.002
    xs, us = np.array([[0, np.pi, 0, 0], [0.5, 2.0, 0, 0]]), np.zeros((2, 200, 1))
    applied = []
    for k in range(5):
        r = O.ddp_solve_batch("cartpole", p, cfg, xs, us, t0=k * 0.004, u_lo=lo, u_hi=hi)
        ua = np.clip(r["u"][:, 0], lo, hi)
        applied.append(ua[:, 0])
        for _ in range(2):
            xs = np.stack([O.model_eval("cartpole", p_sim, 0.0, xs[b], ua[b])["x_next"] for b in range(2)])
        us = r["u"].copy()
    mpc_u = np.array([float(v) for v in d["mpc_u"].split()]).reshape(2, 5)
    np.testing.assert_allclose(mpc_u, np.array(applied).T, rtol=0, atol=1e-8)
    np.testing.assert_allclose(np.array([float(v) for v in d["mpc_x_final"].split()]).reshape(2, 4), xs, rtol=0, atol=1e-9)

    var = {"x": np.zeros((1, 101, 4)), "u": np.zeros((1, 100, 1)), "lambda": np.zeros((1, 101, 4)),
           "s": np.ones((1, 100, 4)), "nu": np.ones((1, 100, 4))}
    fref = O.fmpc_solve_batch("fmpc_cartpole", O.default_params("fmpc_cartpole"), O.fmpc_config(max_iter=5), x0, var)
    assert int(d["fmpc_status"]) == fref["status"][0]
    kkt = np.array([float(v) for v in d["fmpc_kkt"].split()])
    np.testing.assert_allclose(kkt, fref["trace"][0, :fref["n_trace"][0], 1], rtol=1e-8)
    fu = np.array([float(v) for v in d["fmpc_u"].split()])
    assert np.max(np.abs(fu - fref["u"][0, :, 0])) <= 1e-8 * (1 + np.max(np.abs(fu)))
    K0 = np.array([float(v) for v in d["fmpc_K0"].split()])
    np.testing.assert_allclose(K0, fref["K"][0, 0], rtol=1e-7, atol=1e-10)
    assert d["fmpc_invalid_argument"] == "[FMPC] x_list length should be 101 but 100."
    assert "must be non-negative" in d["fmpc_runtime_error"]
