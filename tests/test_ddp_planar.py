"""The control-limited DDP backward pass with TWO inputs (BoxQP<2> inside DDPSolver<6, 2>, DDPSolver.hpp:450-497)
against golden vectors produced by the REFERENCE's unmodified DDPSolver.h/.hpp + BoxQP.h on the planar quadrotor
(tests/golden/make_golden_ddp_planar.py -> reference_ddp_planar.npz; reference control flow on the Eigen shim).  The
reference's own tests reach that code with one input only.

CPU tests pin the oracle (trace, trajectories: 1e-10 relative), GPU tests the CUDA path through the C ABI (iteration
counts and return values exact; u 1e-8 relative; cost 1e-10).

Two of the cases (SPLIT_PRONE) are ones where the reference ALGORITHM is discontinuous in the last bit of its input
once inputs sit on a limit: forwardPass leaves u = limit +- 1 ulp there, the next QP's box is [lo - u, hi - u] and its
warm start is the neighbouring step's k = hi - u', and BoxQP decides its free set -- hence whether that step gets a
feedback gain at all -- by the exact comparison x == upper (BoxQP.h:189-191).  test_reference_splits_on_the_last_bit_
at_a_limit shows the reference's own result moving by O(1) under a half-ulp change of x0; tools/diag_boxqp_split.py
traces one such step on the GPU.  No implementation with another rounding (FMA contraction on the device) can follow
those instances beyond the iteration where that happens, so the CUDA path is held to them exactly for the iterations
before (the *_it2 vectors), and afterwards to: at least half of the instances still on the reference's iterates, all of
them finite and improved."""
import os

import numpy as np
import pytest

import oracle_lib as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_ddp_planar.npz"))
SPLIT_PRONE = ["planar_box_cross_fixed", "planar_box_cold"]
STRICT = ["planar_free", "planar_box_wide", "planar_box_tight", "planar_box_mixed", "planar_box_cross_fixed_it2",
          "planar_box_cold_it2"]
CASES = STRICT + SPLIT_PRONE
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
    for name in ("planar_box_tight", "planar_box_mixed", "planar_box_cross_fixed", "planar_box_cold"):
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


def _perturbed_by_half_an_ulp(x0, seed):
    rng = np.random.default_rng(seed)
    return x0 * (1.0 + 1.2e-16 * rng.choice([-1, 0, 1], size=x0.shape))


def test_reference_splits_on_the_last_bit_at_a_limit():
    """The SPLIT_PRONE cases: the oracle (bit-identical to the reference headers here, see the test above) solved again
    from x0 changed in its last bit ends O(0.1 .. 1) away in u on some instance -- from the third iteration on, never
    in the first two; the STRICT limited cases do not move."""
    def spread(name, max_iter=None):
        c, cfg, lo, hi = _case(name)
        if max_iter is not None:
            cfg["max_iter"] = max_iter
        ui = np.repeat(c["u_init"][None], len(c["x0"]), axis=0)
        run = lambda x0: O.ddp_solve_batch("planar_quadrotor", c["params"], O.ddp_config(**cfg), x0, ui, u_lo=lo, u_hi=hi)
        base = run(c["x0"])
        return np.max([_rel(run(_perturbed_by_half_an_ulp(c["x0"], s))["u"], base["u"]) for s in range(12)], axis=0)

    for name in SPLIT_PRONE:
        assert spread(name).max() > 0.1, name
        assert spread(name, max_iter=2).max() < 1e-12, name
    for name in ("planar_box_tight", "planar_box_mixed"):
        assert spread(name).max() < 1e-12, name


def _gpu_solve(gpu, c, cfg, lo, hi):
    B = len(c["x0"])
    solver = gpu.DDPSolver("planar_quadrotor", params=c["params"], batch_capacity=B)
    for k, v in cfg.items():
        setattr(solver.config(), k, bool(v) if k == "with_input_constraint" else v)
    if lo is not None:
        solver.setInputLimitsFunc((lo, hi))
    ok = solver.solve_batch(0.0, c["x0"], np.repeat(c["u_init"][None], B, axis=0))
    return solver, ok


@pytest.mark.gpu
@pytest.mark.parametrize("name", SPLIT_PRONE)
def test_cuda_on_the_split_prone_cases(gpu, name):
    """Past the iteration where the reference itself becomes last-bit dependent: see the module docstring."""
    c, cfg, lo, hi = _case(name)
    solver, ok = _gpu_solve(gpu, c, cfg, lo, hi)
    u = solver.controlData().u_list
    on_track = _rel(u, c["u"]) <= 1e-8
    assert on_track.mean() >= 0.5, _rel(u, c["u"])
    assert np.array_equal(solver.n_trace()[on_track], c["n_trace"][on_track])
    assert np.array_equal(ok.astype(int)[on_track], c["solve_ret"][on_track])
    cost = c["cost_list"].sum(axis=1)
    assert np.max(np.abs(solver.cost() - cost)[on_track] / np.abs(cost)[on_track]) <= 1e-10
    assert np.all(np.isfinite(u)) and np.all(solver.cost() < solver.trace()[:, 0, 1])
    # the limits hold wherever the feedforward term alone moved the input (forwardPass does not clamp, DDPSolver.hpp:548)
    assert np.all(solver.cost() <= 1.5 * cost)
    solver.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", STRICT)
def test_cuda_matches_the_reference_headers(gpu, name):
    c, cfg, lo, hi = _case(name)
    solver, ok = _gpu_solve(gpu, c, cfg, lo, hi)
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
    # two iterations: beyond, single instances of the reference are last-bit dependent (module docstring)
    kw = dict(horizon_steps=N, max_iter=2, with_input_constraint=1, k_rel_norm_thre=0.0, cost_update_thre=0.0,
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
