"""The control-limited DDP backward pass with TWO inputs (BoxQP<2> inside DDPSolver<6, 2>, DDPSolver.hpp:450-497)
against golden vectors produced by the REFERENCE's unmodified DDPSolver.h/.hpp + BoxQP.h on the planar quadrotor
(tests/golden/make_golden_ddp_planar.py -> reference_ddp_planar.npz; reference control flow on the Eigen shim).  The
reference's own tests reach that code with one input only.

CPU tests pin the oracle (trace, trajectories: 1e-10 relative), GPU tests the CUDA path through the C ABI (iteration
counts and return values exact; u 1e-8 relative; cost 1e-10)."""
import os

import numpy as np
import pytest

import oracle_lib as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_ddp_planar.npz"))
CASES = ["planar_free", "planar_box_wide", "planar_box_tight", "planar_box_mixed", "planar_box_cross_fixed",
         "planar_box_cold"]
INT_KEYS = ("max_iter", "with_input_constraint", "reg_type")


def _rel(a, b):
    ax = tuple(range(1, a.ndim))
    return np.max(np.abs(a - b), axis=ax) / (1.0 + np.max(np.abs(b), axis=ax))


def _case(name):
    c = {k.split("/", 1)[1]: G[k] for k in G.files if k.startswith(name + "/")}
    cfg = {k[4:]: (int(v) if k[4:] in INT_KEYS else float(v)) for k, v in c.items() if k.startswith("cfg_")}
    cfg["horizon_steps"] = int(c["N"])
    lo, hi = (c["limits"][0], c["limits"][1]) if "limits" in c else (None, None)
    return c, cfg, lo, hi


def test_the_cases_clamp_inputs():
    """The vectors exercise what they are for: each limited case has steps with both, one and no input at a limit."""
    for name in CASES[2:]:
        c, _, lo, hi = _case(name)
        at = (c["u"] <= lo + 1e-12) | (c["u"] >= hi - 1e-12)
        n_clamped = at.sum(axis=2)
        assert {0, 1, 2} <= set(np.unique(n_clamped)), (name, np.unique(n_clamped))
    assert 0 in G["planar_box_tight/solve_ret"] and 1 in G["planar_box_tight/solve_ret"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_the_reference_headers(name):
    c, cfg, lo, hi = _case(name)
    B, N = len(c["x0"]), cfg["horizon_steps"]
    r = O.ddp_solve_batch("planar_quadrotor", c["params"], O.ddp_config(**cfg), c["x0"],
                          np.repeat(c["u_init"][None], B, axis=0), u_lo=lo, u_hi=hi)
    assert np.array_equal(r["n_trace"], c["n_trace"])
    assert np.array_equal((r["status"] == 1).astype(int), c["solve_ret"])
    np.testing.assert_allclose(r["trace"], c["trace"], rtol=1e-9, atol=1e-12)
    assert _rel(r["u"], c["u"]).max() <= 1e-10
    assert _rel(r["x"], c["x"]).max() <= 1e-10
    np.testing.assert_allclose(r["cost_list"], c["cost_list"], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_the_reference_headers(gpu, name):
    c, cfg, lo, hi = _case(name)
    B, N = len(c["x0"]), cfg["horizon_steps"]
    solver = gpu.DDPSolver("planar_quadrotor", params=c["params"], batch_capacity=B)
    for k, v in cfg.items():
        setattr(solver.config(), k, bool(v) if k == "with_input_constraint" else v)
    if lo is not None:
        solver.setInputLimitsFunc((lo, hi))
    ok = solver.solve_batch(0.0, c["x0"], np.repeat(c["u_init"][None], B, axis=0))
    assert np.array_equal(solver.n_trace(), c["n_trace"])
    assert np.array_equal(ok.astype(int), c["solve_ret"])
    assert _rel(solver.controlData().u_list, c["u"]).max() <= 1e-8
    assert _rel(solver.controlData().x_list, c["x"]).max() <= 1e-8
    cost = c["cost_list"].sum(axis=1)
    assert np.max(np.abs(solver.cost() - cost) / np.abs(cost)) <= 1e-10
    tr = solver.trace()
    np.testing.assert_array_equal(tr[:, :, 0], c["trace"][:, :, 0])
    np.testing.assert_allclose(tr[:, :, 1:5], c["trace"][:, :, 1:5], rtol=1e-8, atol=1e-300)
    solver.close()


@pytest.mark.gpu
def test_planar_batch_with_limits_against_oracle(gpu):
    """A batch of random starts with per-input limits: CUDA against the oracle (which the test above pins)."""
    B, N = 256, 60
    rng = np.random.default_rng(21)
    x0 = np.concatenate([rng.uniform(-1.5, 1.5, (B, 2)), rng.uniform(-0.5, 0.5, (B, 1)), rng.uniform(-0.8, 0.8, (B, 3))],
                        axis=1)
    p = O.default_params("planar_quadrotor")
    hover = 0.5 * p[1] * 9.80665
    lo, hi = np.array([0.85 * hover, 0.5 * hover]), np.array([1.25 * hover, 1.1 * hover])
    u_init = np.full((B, N, 2), hover)
    kw = dict(horizon_steps=N, max_iter=6, with_input_constraint=1, k_rel_norm_thre=0.0, cost_update_thre=0.0,
              cost_update_ratio_thre=0.0, lambda_thre=0.0)
    ref = O.ddp_solve_batch("planar_quadrotor", p, O.ddp_config(**kw), x0, u_init, u_lo=lo, u_hi=hi)
    solver = gpu.DDPSolver("planar_quadrotor", params=p, batch_capacity=B)
    for k, v in kw.items():
        setattr(solver.config(), k, bool(v) if k == "with_input_constraint" else v)
    solver.setInputLimitsFunc((lo, hi))
    solver.solve_batch(0.0, x0, u_init)
    assert np.array_equal(solver.n_trace(), ref["n_trace"])
    assert _rel(solver.controlData().u_list, ref["u"]).max() <= 1e-8
    np.testing.assert_allclose(solver.cost(), ref["cost_list"].sum(axis=1), rtol=1e-10)
    solver.close()
