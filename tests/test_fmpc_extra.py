"""The two FMPC code paths none of the reference's own tests reach, against golden vectors produced by the REFERENCE's
unmodified FmpcSolver.h/.hpp (tests/golden/make_golden_fmpc_extra.py -> reference_fmpc_extra.npz):

  * two inputs (planar quadrotor, FmpcSolver<6, 2, 4>): Eigen::LDLT of G with diagonal pivoting, the Eigen::FullPivLU
    fallback when LDLT reports NumericalIssue and the break_if_llt_fails exit (FmpcSolver.hpp:596-617);
  * a time-varying inequality dimension (windowed cart-pole, FmpcSolver<4, 1, Eigen::Dynamic>, FmpcProblem.h:62-86).

CPU tests pin the oracle, GPU tests the CUDA path (through the C ABI): iterates (x, u, lambda, s, nu) after 1 .. 10
iterations, relative 1e-8, the KKT sequence 1e-7 and the status words."""
import os

import numpy as np
import pytest

import oracle_lib as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_fmpc_extra.npz"))
REL_TOL = 1e-8
PLANAR = ["planar_it1", "planar_it2", "planar_it3", "planar_it5", "planar_it10", "planar_fullpivlu", "planar_break"]
WINDOWED = ["windowed_it1", "windowed_it3", "windowed_it10", "windowed_t0", "windowed_initcomp", "windowed_linesearch"]


def _rel(a, b):
    ax = tuple(range(1, a.ndim))
    return np.max(np.abs(a - b), axis=ax) / (1.0 + np.max(np.abs(b), axis=ax))


def _case(name):
    c = {k.split("/", 1)[1]: G[k] for k in G.files if k.startswith(name + "/")}
    cfg = {k[4:]: (float(v) if k == "cfg_kkt_error_thre" else int(v)) for k, v in c.items() if k.startswith("cfg_")}
    var = {k[4:]: c[k] for k in c if k.startswith("var_")}
    return c, cfg, var


def _check(name, got, c, n_cmp=None):
    """got: dict with x, u, lambda, s, nu [B, ...], kkt [B, max_iter], status [B]."""
    assert np.array_equal(got["status"], c["status"]), (name, got["status"], c["status"])
    if int(c["status"][0]) == 3:
        return  # ErrorInBackward: the reference leaves the sweep early, iterates are not comparable
    for key in ("x", "u", "lambda", "s", "nu"):
        err = _rel(got[key], c[key]).max()
        assert err <= REL_TOL, (name, key, err)
    n = int(c["n_trace"][0])
    np.testing.assert_allclose(got["kkt"][:, :n], c["kkt"][:, :n], rtol=1e-7, err_msg=name)


@pytest.mark.parametrize("name", PLANAR + WINDOWED)
def test_oracle_matches_the_reference_headers(name):
    model = "fmpc_planar_quadrotor" if name.startswith("planar") else "fmpc_cartpole_windowed"
    c, cfg, var = _case(name)
    B = len(c["x0"])
    vb = {k: np.repeat(v[None], B, axis=0) for k, v in var.items()}
    out = O.fmpc_solve_batch(model, c["params"], O.fmpc_config(**cfg), c["x0"], vb, t0=float(c["t0"]))
    got = {k: out[k] for k in ("x", "u", "lambda", "s", "nu", "status")}
    got["kkt"] = out["trace"][:, :, 1]
    _check(name, got, c)


def test_windowed_dimension_sequence():
    """The test problem really changes its dimension inside the horizon, and the padding rows come back neutral."""
    c, cfg, _ = _case("windowed_it3")
    p, dt, t0 = c["params"], c["params"][0], float(c["t0"])
    dims = np.array([4 if (p[14] <= t0 + i * dt < p[15]) else 2 for i in range(cfg["horizon_steps"])])
    assert set(dims) == {2, 4} and dims[0] == 2 and dims[-1] == 2
    assert np.all(c["s"][:, dims == 2, 2:] == 1.0) and np.all(c["nu"][:, dims == 2, 2:] == 0.0)
    assert np.all(c["nu"][:, dims == 4, 2:] > 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", PLANAR + WINDOWED)
def test_cuda_matches_the_reference_headers(gpu, name):
    model = "planar_quadrotor" if name.startswith("planar") else "cartpole_windowed"
    c, cfg, var = _case(name)
    B = len(c["x0"])
    s = gpu.FmpcSolver(model, params=c["params"], batch_capacity=B)
    for k, v in cfg.items():
        setattr(s.config(), k, bool(v) if k in ("break_if_llt_fails", "init_complementary_variable", "enable_line_search",
                                                 "check_nan", "update_barrier_eps") else v)
    v = s.make_variable(B)
    for name_v, key in (("x_list", "x"), ("u_list", "u"), ("lambda_list", "lambda"), ("s_list", "s"), ("nu_list", "nu")):
        setattr(v, name_v, np.repeat(var[key][None], B, axis=0).copy())
    status = s.solve_batch(float(c["t0"]), c["x0"], v)
    out = s.variable()
    got = {"x": out.x_list, "u": out.u_list, "lambda": out.lambda_list, "s": out.s_list, "nu": out.nu_list,
           "status": np.asarray(status), "kkt": s.trace()[:, :, 1]}
    _check(name, got, c)
    s.close()


@pytest.mark.gpu
def test_planar_quadrotor_batch_against_oracle(gpu):
    """A batch of random starts through 6 iterations: CUDA against the oracle (which the test above pins)."""
    B, N = 128, 60
    rng = np.random.default_rng(9)
    x0 = np.concatenate([rng.uniform(-1, 1, (B, 2)), rng.uniform(-0.4, 0.4, (B, 1)), rng.uniform(-0.5, 0.5, (B, 3))], axis=1)
    p = O.default_params("fmpc_planar_quadrotor")
    hover = 0.5 * p[1] * 9.80665
    s = gpu.FmpcSolver("planar_quadrotor", params=p, batch_capacity=B)
    s.config().horizon_steps, s.config().max_iter = N, 6
    v = s.make_variable(B)
    v.reset(0.0, hover, 0.0, 1.0, 1.0)
    status = s.solve_batch(0.0, x0, v)
    ref = O.fmpc_solve_batch("fmpc_planar_quadrotor", p, O.fmpc_config(horizon_steps=N, max_iter=6), x0,
                             {"x": v.x_list, "u": v.u_list, "lambda": v.lambda_list, "s": v.s_list, "nu": v.nu_list})
    assert np.array_equal(np.asarray(status), ref["status"])
    out = s.variable()
    for key, arr in (("x", out.x_list), ("u", out.u_list), ("lambda", out.lambda_list), ("s", out.s_list),
                     ("nu", out.nu_list)):
        assert _rel(arr, ref[key]).max() <= REL_TOL, key
    s.close()
