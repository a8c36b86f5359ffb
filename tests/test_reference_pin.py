"""Pins against outputs of the REFERENCE's own code.

tests/golden/reference_outputs.npz was produced by tests/golden/make_golden.py from oracle/_ref =
/root/reference's DDPSolver.h(.hpp), BoxQP.h, FmpcSolver.h(.hpp) compiled UNMODIFIED against the
Eigen-subset shim (oracle/ref/eigen_shim; the image has no Eigen).  Three things are checked against it:
the CPU oracle (restatement), the shim build itself when the reference checkout is present, and -- on the
GPU box -- the CUDA path through the C ABI."""
import os

import numpy as np
import pytest

import oracle_lib as O
import ref_lib as R

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz"))
DDP_CASES = ["ddp_ref", "ddp_fixed", "ddp_reg2", "ddp_default_500", "ddp_box200", "ddp_box400", "ddp_box_tight"]
KATS_GT = [[-1.5, -2.0], [0.5, -2.0], [0.0, -1.0], [-2.0, -1.0], [-2.0, -3.0]]


def _ddp_cfg(name):
    max_iter, box, k_thre, c_thre, reg = GOLDEN[f"{name}/cfg"]
    return dict(horizon_steps=int(GOLDEN[f"{name}/N"]), max_iter=int(max_iter), with_input_constraint=int(box),
                k_rel_norm_thre=float(k_thre), cost_update_thre=float(c_thre), reg_type=int(reg))


def _limits(name):
    if f"{name}/limits" in GOLDEN:
        lo, hi = GOLDEN[f"{name}/limits"]
        return np.array([lo]), np.array([hi])
    return None, None


def _rel(a, b):
    ax = tuple(range(1, a.ndim))
    return np.max(np.abs(a - b), axis=ax) / (1.0 + np.max(np.abs(b), axis=ax))


def test_reference_boxqp_known_answers():
    """The reference's BoxQP<2> and BoxQP<Dynamic> (run through the shim) on TestBoxQP.cpp:39-98: tol 1e-6."""
    x = GOLDEN["boxqp/x"]
    for dyn in range(2):
        for i, gt in enumerate(KATS_GT):
            assert np.linalg.norm(x[dyn * 5 + i] - np.array(gt)) < 1e-6
    H = np.array([[1.0, 0.0], [0.0, 0.5]])
    kats = [([1.5, 1.0], [-10, -10], [10, 10]), ([1.5, 1.0], [0.5, -2.0], [5.0, 2.0]),
            ([1.0, 1.5], [0.0, -1.0], [5.0, -0.5]), ([1.5, 1.0], [-5.0, -1.0], [-2.0, 2.0]),
            ([1.0, 1.5], [-5.0, -10.0], [-2.0, 10.0])]
    for i, (g, lo, hi) in enumerate(kats):  # the oracle agrees with the reference code bit for bit, exit code included
        xo, rv, _ = O.boxqp_solve(H, g, lo, hi)
        np.testing.assert_array_equal(xo, x[i])
        assert rv == GOLDEN["boxqp/retval"][i]


@pytest.mark.parametrize("name", DDP_CASES)
def test_oracle_matches_reference_ddp(name):
    cfg = O.ddp_config(**_ddp_cfg(name))
    x0 = GOLDEN[f"{name}/x0"]
    B, N = x0.shape[0], cfg.horizon_steps
    lo, hi = _limits(name)
    r = O.ddp_solve_batch("cartpole", O.default_params("cartpole"), cfg, x0, np.zeros((B, N, 1)), u_lo=lo, u_hi=hi)
    assert np.array_equal(r["n_trace"], GOLDEN[f"{name}/n_trace"])
    assert np.array_equal((r["status"] == 1).astype(int), GOLDEN[f"{name}/solve_ret"])
    np.testing.assert_allclose(r["trace"], GOLDEN[f"{name}/trace"], rtol=1e-11, atol=1e-13)
    assert _rel(r["u"], GOLDEN[f"{name}/u"]).max() <= 1e-11
    assert _rel(r["x"], GOLDEN[f"{name}/x"]).max() <= 1e-11
    np.testing.assert_allclose(r["cost_list"], GOLDEN[f"{name}/cost_list"], rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("name,model", [("fmpc_cartpole", "fmpc_cartpole"), ("fmpc_oscillator", "fmpc_oscillator")])
def test_oracle_matches_reference_fmpc(name, model):
    nx, nu, ng, _ = O.model_dims(model)
    x0 = GOLDEN[f"{name}/x0"]
    B, N, max_iter = x0.shape[0], int(GOLDEN[f"{name}/N"]), int(GOLDEN[f"{name}/max_iter"])
    var = {"x": np.zeros((B, N + 1, nx)), "u": np.zeros((B, N, nu)), "lambda": np.zeros((B, N + 1, nx)),
           "s": np.ones((B, N, ng)), "nu": np.ones((B, N, ng))}
    r = O.fmpc_solve_batch(model, O.default_params(model), O.fmpc_config(horizon_steps=N, max_iter=max_iter), x0, var)
    assert np.array_equal(r["status"], GOLDEN[f"{name}/status"])
    assert np.array_equal(r["n_trace"], GOLDEN[f"{name}/n_trace"])
    np.testing.assert_allclose(r["trace"][:, :, 1], GOLDEN[f"{name}/kkt"], rtol=1e-9)
    for key in ("x", "u", "lambda", "s", "nu"):
        assert _rel(r[key], GOLDEN[f"{name}/{key}"]).max() <= 1e-9, key
    assert _rel(r["K"], GOLDEN[f"{name}/K"]).max() <= 1e-8


def test_fmpc_cartpole_kkt_sequence_known():
    """SURVEY.md App. C probe for x0=(0,pi,0,0): the reference code gives the same KKT-error sequence."""
    want = [334.4, 171.0, 571.4, 1313.5, 894.2, 788.4, 1031.0, 658.4, 15.06, 7.13]
    np.testing.assert_allclose(GOLDEN["fmpc_cartpole/kkt"][0], want, rtol=2e-3)
    assert GOLDEN["fmpc_cartpole/status"][0] == 5  # MaxIterationReached


@pytest.mark.skipif(not R.available(), reason="needs the reference checkout (/root/reference)")
def test_golden_file_is_what_the_reference_code_produces():
    """Re-run the reference headers (shim build) and compare with the committed vectors."""
    p = O.default_params("cartpole")
    for name in ("ddp_ref", "ddp_box200"):
        cfg = R.ddp_config(**_ddp_cfg(name))
        lo, hi = _limits(name)
        for i, x0 in enumerate(GOLDEN[f"{name}/x0"]):
            o = R.ddp_solve_cartpole(p, cfg, x0, np.zeros(cfg.horizon_steps), u_lo=lo, u_hi=hi)
            np.testing.assert_array_equal(o["u"], GOLDEN[f"{name}/u"][i])
            np.testing.assert_array_equal(o["trace"], GOLDEN[f"{name}/trace"][i])
    # the reference's Configuration constructor and the C ABI / oracle defaults agree field by field
    a, b = R.ddp_config(), O.ddp_config()
    assert bytes(a) == bytes(b)


@pytest.mark.skipif(not os.environ.get("NMPC_REF_EIGEN_LIB"),
                    reason="needs the reference headers built against a REAL Eigen (make -C oracle/ref eigen EIGEN_DIR=...; "
                           "NMPC_REF_EIGEN_LIB=<the .so>): not available in this image")
def test_golden_vectors_against_a_real_eigen_build():
    """Where Eigen exists: the committed vectors (reference control flow on the Eigen shim) against the same reference
    code on real Eigen.  Tolerances are the ones the oracle and the CUDA path are held to, not bit-exactness: Eigen's
    packet reductions and blocked products round differently from the shim's k-ascending sums."""
    p = O.default_params("cartpole")
    for name in ("ddp_ref", "ddp_box200", "ddp_reg2"):
        cfg = R.ddp_config(**_ddp_cfg(name))
        lo, hi = _limits(name)
        for i, x0 in enumerate(GOLDEN[f"{name}/x0"]):
            o = R.ddp_solve_cartpole(p, cfg, x0, np.zeros(cfg.horizon_steps), u_lo=lo, u_hi=hi)
            assert o["n_trace"] == GOLDEN[f"{name}/n_trace"][i]
            assert _rel(o["u"][None], GOLDEN[f"{name}/u"][i][None]).max() <= 1e-9
            np.testing.assert_allclose(o["trace"][:, 1], GOLDEN[f"{name}/trace"][i][:, 1], rtol=1e-9)
    for name, model in (("fmpc_cartpole", "fmpc_cartpole"), ("fmpc_oscillator", "fmpc_oscillator")):
        nx, nu, ng, _ = O.model_dims(model)
        N, max_iter = int(GOLDEN[f"{name}/N"]), int(GOLDEN[f"{name}/max_iter"])
        var = {"x": np.zeros((N + 1, nx)), "u": np.zeros((N, nu)), "lambda": np.zeros((N + 1, nx)), "s": np.ones((N, ng)),
               "nu": np.ones((N, ng))}
        cfg = O.fmpc_config(horizon_steps=N, max_iter=max_iter)
        for i, x0 in enumerate(GOLDEN[f"{name}/x0"]):
            o = R.fmpc_solve(model, O.default_params(model), cfg, x0, var)
            assert o["status"] == GOLDEN[f"{name}/status"][i]
            for key in ("x", "u", "lambda", "s", "nu"):
                assert _rel(o[key][None], GOLDEN[f"{name}/{key}"][i][None]).max() <= 1e-8, (name, key)


@pytest.mark.gpu
@pytest.mark.parametrize("name", DDP_CASES)
def test_cuda_matches_reference_ddp(gpu, name):
    kw = _ddp_cfg(name)
    x0 = GOLDEN[f"{name}/x0"]
    B, N = x0.shape[0], kw["horizon_steps"]
    solver = gpu.DDPSolver("cartpole", batch_capacity=B)
    c = solver.config()
    for k, v in kw.items():
        setattr(c, k, bool(v) if k == "with_input_constraint" else v)
    lo, hi = _limits(name)
    if lo is not None:
        solver.setInputLimitsFunc((lo, hi))
    ok = solver.solve_batch(0.0, x0, np.zeros((B, N, 1)))
    fixed = kw["k_rel_norm_thre"] == 0.0
    assert np.array_equal(solver.n_trace(), GOLDEN[f"{name}/n_trace"])
    assert np.array_equal(ok.astype(int), GOLDEN[f"{name}/solve_ret"])
    u_tol, c_tol = (1e-6, 1e-10) if fixed else (1e-9, 1e-12)
    assert _rel(solver.controlData().u_list, GOLDEN[f"{name}/u"]).max() <= u_tol
    cost = GOLDEN[f"{name}/cost_list"].sum(axis=1)
    assert np.max(np.abs(solver.cost() - cost) / np.abs(cost)) <= c_tol
    tr = solver.trace()
    np.testing.assert_array_equal(tr[:, :, 0], GOLDEN[f"{name}/trace"][:, :, 0])
    if not fixed:
        np.testing.assert_allclose(tr[:, :, 1:5], GOLDEN[f"{name}/trace"][:, :, 1:5], rtol=1e-10, atol=1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("name,model", [("fmpc_cartpole", "cartpole"), ("fmpc_oscillator", "oscillator")])
def test_cuda_matches_reference_fmpc(gpu, name, model):
    x0 = GOLDEN[f"{name}/x0"]
    B, N, max_iter = x0.shape[0], int(GOLDEN[f"{name}/N"]), int(GOLDEN[f"{name}/max_iter"])
    solver = gpu.FmpcSolver(model, batch_capacity=B)
    c = solver.config()
    c.horizon_steps, c.max_iter = N, max_iter
    var = solver.make_variable(B)
    var.reset(0.0, 0.0, 0.0, 1.0, 1.0)
    status = solver.solve_batch(0.0, x0, var)
    assert np.array_equal(status, GOLDEN[f"{name}/status"])
    v = solver.variable()
    out = {"x": v.x_list, "u": v.u_list, "lambda": v.lambda_list, "s": v.s_list, "nu": v.nu_list}
    for key in out:
        assert _rel(out[key], GOLDEN[f"{name}/{key}"]).max() <= 1e-8, key
    np.testing.assert_allclose(solver.trace()[:, :, 1], GOLDEN[f"{name}/kkt"], rtol=1e-7)
