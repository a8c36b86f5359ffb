"""GPU parity tests: the CUDA DDP path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances (fp64), SURVEY.md 8(d) / BASELINE.md 5:
  M-ref   (reference termination): max|du| / (1 + max|u|) <= 1e-9, |dcost|/|cost| <= 1e-12,
          identical iteration count and status per instance.
  M-fixed (10 forced iterations): |dcost|/|cost| <= 1e-10, relative du <= 1e-6.
"""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

U_TOL_REF = 1e-9
COST_TOL_REF = 1e-12
U_TOL_FIXED = 1e-6
COST_TOL_FIXED = 1e-10


def _rel_u(a, b):
    return np.max(np.abs(a - b), axis=(1, 2)) / (1.0 + np.max(np.abs(b), axis=(1, 2)))


def _solve_both(gpu, B, seed, N=100, **cfg_kw):
    p = O.default_params("cartpole")
    x0 = O.cartpole_x0(B, seed)
    u_init = np.zeros((B, N, 1))
    ocfg = O.ddp_config(horizon_steps=N, **cfg_kw)
    ref = O.ddp_solve_batch("cartpole", p, ocfg, x0, u_init)
    solver = gpu.DDPSolver("cartpole", params=p, batch_capacity=B)
    c = solver.config()
    c.horizon_steps = N
    for k, v in cfg_kw.items():
        setattr(c, k, v)
    ok = solver.solve_batch(0.0, x0, u_init)
    return ref, solver, ok


def test_model_functor_matches_oracle(gpu):
    """Device functor vs the oracle's independent restatement, incl. the reference's check point
    (TestDDPCartPole.cpp:620-623) and its central-difference derivative check (:629-648)."""
    rng = np.random.default_rng(7)
    n = 64
    x = rng.uniform(-3, 3, (n, 4))
    x[0] = [1.0, -2.0, 3.0, -4.0]
    u = rng.uniform(-20, 20, (n, 1))
    u[0] = 10.0
    p = O.default_params("cartpole")
    d = gpu.model_eval("cartpole", 0.0, x, u, params=p)
    for i in range(n):
        o = O.model_eval("cartpole", p, 0.0, x[i], u[i])
        for key in ("x_next", "Fx", "Fu", "Lx", "Lu", "Lxx", "Luu", "Lxu", "Vx", "Vxx"):
            np.testing.assert_allclose(d[key][i], o[key], rtol=1e-12, atol=1e-13, err_msg=key)
        assert abs(d["running_cost"][i] - o["running_cost"]) <= 1e-12 * max(1, abs(o["running_cost"]))
        assert abs(d["terminal_cost"][i] - o["terminal_cost"]) <= 1e-12 * max(1, abs(o["terminal_cost"]))
    eps = 1e-6
    Fx = np.zeros((4, 4))
    for j in range(4):
        e = np.zeros(4)
        e[j] = eps
        xp = gpu.model_eval("cartpole", 0.0, (x[0] + e)[None], u[:1], params=p)["x_next"][0]
        xm = gpu.model_eval("cartpole", 0.0, (x[0] - e)[None], u[:1], params=p)["x_next"][0]
        Fx[:, j] = (xp - xm) / (2 * eps)
    assert np.linalg.norm(d["Fx"][0] - Fx) < 1e-6


def test_branch_free_device_math(gpu):
    """models/cartpole.h: the functor written for instruction latency (CartPole<double, true>, what the lanes / split
    kernels evaluate) gives the values of the library-math functor -- identical in 98 % of the samples, within 2 ulp in
    the others (the two instantiations fuse multiply-adds differently) -- over the range a rollout visits, and NaN, not
    a wrong number, beyond |theta| = 2^31."""
    rng = np.random.default_rng(11)
    n = 4096
    x = rng.uniform(-3, 3, (n, 4))
    x[:, 1] = np.concatenate([rng.uniform(-10, 10, n // 2), rng.uniform(-1e6, 1e6, n // 4), rng.uniform(-2e9, 2e9, n // 4)])
    x[0, 1], x[1, 1], x[2, 1] = 0.0, np.pi, -np.pi / 2
    u = rng.uniform(-20, 20, (n, 1))
    a = gpu.model_eval("cartpole", 0.0, x, u)
    b = gpu.model_eval("cartpole_branch_free", 0.0, x, u)
    for key in ("x_next", "Fx", "Fu", "Lx", "Lu", "running_cost"):
        bad = np.nonzero(np.any((a[key] != b[key]).reshape(n, -1), axis=1))[0]
        print(key, "rows that differ:", bad.size, "max abs diff", np.abs(a[key] - b[key]).max())
        assert bad.size <= n // 20, (key, bad.size)
        np.testing.assert_allclose(a[key], b[key], rtol=5e-16, atol=5e-16 * np.abs(b[key]).max(), err_msg=key)
    far = gpu.model_eval("cartpole_branch_free", 0.0, np.array([[0.0, 3e9, 0.0, 0.0]]), np.array([[1.0]]))
    assert np.all(np.isnan(far["x_next"][0, 2:]))
    with pytest.raises(gpu.NmpcB200Error):
        gpu.DDPSolver("cartpole_branch_free", batch_capacity=1)  # evaluation only: no solver kernels behind that name


def test_tuning_knobs(gpu):
    """nmpc_b200_ddp_set_tuning / _get_tuning: defaults scale with the device's SM count, a pinned variant solves the same
    problem (to rounding), unknown keys are refused."""
    import torch

    sms = torch.cuda.get_device_properties(0).multi_processor_count
    B, N = 512, 60
    x0 = O.cartpole_x0(B, 31)
    s = gpu.DDPSolver("cartpole", batch_capacity=B)
    s.config().horizon_steps, s.config().max_iter = N, 6
    assert s.get_tuning("backward_lanes_max_batch") == sms * 32
    assert s.get_tuning("forward_split_max_batch") == sms * 83 and s.get_tuning("forward_phased_max_batch") == sms * 332
    s.solve_batch(0.0, x0, np.zeros((B, N, 1)))
    u_default, cost_default = s.controlData().u_list, s.cost()
    for knobs in (dict(backward_lanes=0), dict(backward_lanes=2), dict(backward_fused=0), dict(forward_split=0),
                  dict(forward_lanes=1), dict(backward_lanes_max_batch=256, forward_split_max_batch=256)):
        t = gpu.DDPSolver("cartpole", batch_capacity=B)
        t.config().horizon_steps, t.config().max_iter = N, 6
        t.set_tuning(**knobs)
        assert all(t.get_tuning(k) == v for k, v in knobs.items())
        t.solve_batch(0.0, x0, np.zeros((B, N, 1)))
        assert np.array_equal(t.iterations(), s.iterations()), knobs
        np.testing.assert_allclose(t.cost(), cost_default, rtol=1e-10, err_msg=str(knobs))
        assert np.max(np.abs(t.controlData().u_list - u_default)) <= 1e-7 * (1 + np.max(np.abs(u_default))), knobs
        t.close()
    with pytest.raises(gpu.NmpcB200Error):
        s.set_tuning(no_such_knob=1)
    s.close()


def test_single_instance_swingup(gpu):
    """solve() of one instance: x0=(0,pi,0,0), N=100, max_iter=10 (oracle values pinned in test_oracle_ddp)."""
    solver = gpu.DDPSolver("cartpole", batch_capacity=1)
    solver.config().max_iter = 10
    converged = solver.solve(0.0, [0, np.pi, 0, 0], np.zeros((100, 1)))
    assert converged is True
    tl = solver.traceDataList()
    assert [t.iter for t in tl] == list(range(8))
    want = [498.415022255, 406.088004263, 404.897899816, 404.865142163, 404.86399135, 404.863948982, 404.863947391,
            404.863947331]
    np.testing.assert_allclose([t.cost for t in tl], want, rtol=2e-12)
    assert all(t.alpha == 1.0 for t in tl[1:])
    cd = solver.controlData()
    np.testing.assert_allclose(cd.u_list[0, :4, 0],
                               [18.786169762863928, 18.484618096972063, 18.184692098150514, 17.886397315575014],
                               rtol=1e-10)
    assert abs(cd.cost_list[0].sum() - tl[-1].cost) < 1e-9


@pytest.mark.parametrize("B,seed", [(1, 11), (37, 5), (256, 0), (1000, 3)])
def test_parity_reference_termination(gpu, B, seed):
    """M-ref: max_iter=10 with the reference's termination rules; ragged batch sizes."""
    ref, solver, ok = _solve_both(gpu, B, seed, max_iter=10)
    assert np.array_equal(solver.status(), ref["status"])
    assert np.array_equal(solver.iterations(), ref["iters"])
    assert np.array_equal(solver.n_forward(), ref["n_fwd"])
    assert np.array_equal(solver.n_backward(), ref["n_bwd"])
    assert np.array_equal(ok, ref["status"] == 1)
    cd = solver.controlData()
    assert _rel_u(cd.u_list, ref["u"]).max() <= U_TOL_REF
    assert _rel_u(cd.x_list, ref["x"]).max() <= 1e-8
    cost = solver.cost()
    assert np.max(np.abs(cost - ref["cost"]) / np.abs(ref["cost"])) <= COST_TOL_REF
    assert np.max(np.abs(cd.cost_list.sum(axis=1) - cost) / np.abs(cost)) <= 1e-13
    # gains and the full trace table
    assert _rel_u(solver.k_list(), ref["k"]).max() <= 1e-6
    tr = solver.trace()
    assert np.array_equal(solver.n_trace(), ref["n_trace"])
    np.testing.assert_array_equal(tr[:, :, 0], ref["trace"][:, :, 0])
    np.testing.assert_allclose(tr[:, :, 1:5], ref["trace"][:, :, 1:5], rtol=1e-11, atol=1e-300)
    np.testing.assert_array_equal(solver.u0(), cd.u_list[:, 0, :])


def test_parity_fixed_iterations(gpu):
    """M-fixed: k_rel_norm_thre = cost_update_thre = 0 forces exactly 10 iterations (roofline mode)."""
    ref, solver, _ = _solve_both(gpu, 512, 0, max_iter=10, k_rel_norm_thre=0.0, cost_update_thre=0.0)
    assert np.all(ref["iters"] == 10)
    assert np.array_equal(solver.iterations(), ref["iters"])
    cost = solver.cost()
    assert np.max(np.abs(cost - ref["cost"]) / np.abs(ref["cost"])) <= COST_TOL_FIXED
    assert _rel_u(solver.controlData().u_list, ref["u"]).max() <= U_TOL_FIXED


def test_parity_longer_horizon_and_reg_type2(gpu):
    ref, solver, _ = _solve_both(gpu, 64, 9, N=200, max_iter=8, reg_type=2)
    assert np.array_equal(solver.iterations(), ref["iters"])
    assert np.array_equal(solver.status(), ref["status"])
    assert _rel_u(solver.controlData().u_list, ref["u"]).max() <= U_TOL_REF
    assert np.max(np.abs(solver.cost() - ref["cost"]) / np.abs(ref["cost"])) <= COST_TOL_REF


def test_lambda_retry_and_failure_paths(gpu):
    """A negative input weight makes Quu indefinite: the LLT failure rule (pivot <= 0) must drive the
    same lambda-increase retries, and lambda_max the same failures, as in the oracle."""
    p = O.default_params("cartpole")
    p[8] = -5e-4  # running_u < 0
    B, N = 128, 60
    x0 = O.cartpole_x0(B, 21)
    u_init = np.zeros((B, N, 1))
    for lam_max in (1e10, 1e-3):
        ocfg = O.ddp_config(horizon_steps=N, max_iter=6, lambda_max=lam_max)
        ref = O.ddp_solve_batch("cartpole", p, ocfg, x0, u_init)
        assert (ref["n_bwd"] > ref["iters"]).any(), "test input does not exercise the retry path"
        solver = gpu.DDPSolver("cartpole", params=p, batch_capacity=B)
        c = solver.config()
        c.horizon_steps, c.max_iter, c.lambda_max = N, 6, lam_max
        solver.solve_batch(0.0, x0, u_init)
        assert np.array_equal(solver.status(), ref["status"])
        assert np.array_equal(solver.n_backward(), ref["n_bwd"])
        assert np.array_equal(solver.iterations(), ref["iters"])
        np.testing.assert_allclose(solver.trace()[:, :, 2], ref["trace"][:, :, 2], rtol=1e-12)
        if lam_max < 1:
            assert (ref["status"] == -1).any()


def test_argument_errors_mirror_reference(gpu):
    solver = gpu.DDPSolver("cartpole", batch_capacity=8)
    solver.config().max_iter = 2
    # DDPSolver.hpp:41-45 std::invalid_argument("initial_u_list length should be 100 but 99.")
    with pytest.raises(ValueError) as e:
        solver.solve(0.0, np.zeros(4), np.zeros((99, 1)))
    assert "initial_u_list length should be 100 but 99." in str(e.value)
    # DDPSolver.hpp:391-414: second-order dynamics derivatives throw std::runtime_error
    solver.config().use_state_eq_second_derivative = True
    with pytest.raises(gpu.NmpcB200Error) as e:
        solver.solve(0.0, np.zeros(4), np.zeros((100, 1)))
    assert e.value.code == 2 and "Vector-tensor product is not implemented yet." in e.value.message
    solver.config().use_state_eq_second_derivative = False
    with pytest.raises(gpu.NmpcB200Error) as e:
        solver.solve_batch(0.0, np.zeros((9, 4)), np.zeros((9, 100, 1)))
    assert e.value.code == 6
    # max_iter exhausted without convergence => solve() returns false (DDPSolver.hpp:140)
    assert solver.solve(0.0, [0, np.pi, 0, 0], np.zeros((100, 1))) is False
    assert solver.status()[0] == 0 and solver.iterations()[0] == 2


def test_device_resident_io_and_warm_start(gpu):
    """Device pointers in/out (torch CUDA tensors) give the same result as host arrays; a second solve
    warm-started with the previous u_list converges immediately, like the MPC loops of the reference
    (TestDDPCartPole.cpp:388-396)."""
    import torch

    B, N = 200, 100
    x0 = O.cartpole_x0(B, 2)
    solver = gpu.DDPSolver("cartpole", batch_capacity=B)
    solver.config().max_iter = 10
    solver.solve_batch(0.0, x0, np.zeros((B, N, 1)))
    u_host = solver.controlData().u_list
    xd = torch.from_numpy(x0).cuda()
    ud = torch.zeros((B, N, 1), dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        solver.solve_batch(0.0, xd, ud, stream=st, read_status=False)
        out = torch.empty((B, N, 1), dtype=torch.float64, device="cuda")
        solver.get_into(1, out, stream=st)
    st.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), u_host)
    it0 = solver.iterations().copy()
    solver.solve_batch(0.0, x0, u_host)
    assert solver.iterations().mean() < it0.mean() and solver.iterations().max() <= 3


def test_large_max_iter_early_exit_matches(gpu):
    """Default max_iter=500: the host polls for 'all finished'; results equal the oracle's."""
    ref, solver, _ = _solve_both(gpu, 96, 4, max_iter=500)
    assert np.array_equal(solver.iterations(), ref["iters"])
    assert np.array_equal(solver.status(), ref["status"])
    assert _rel_u(solver.controlData().u_list, ref["u"]).max() <= U_TOL_REF
    tr = solver.trace()
    assert tr.shape == (96, 501, 9)
    assert np.all(tr[np.arange(96), ref["n_trace"] - 1, 0] == ref["iters"])


@pytest.mark.parametrize("N", [200, 400])
def test_constrained_config1_first_tick(gpu, N):
    """BASELINE.json configs[0]: TestDDPCartPole (with_input_constraint, +-15 N, max_iter 3, x0=(0,pi,0,0));
    N=200 is what the reference's rostest runs (TestDDPCartPole.test:14), N=400 the config's literal text.
    BoxQP-constrained backward pass on the device vs the oracle (whose BoxQP passes the reference's KATs)."""
    p = O.default_params("cartpole")
    lo, hi = np.array([-15.0]), np.array([15.0])
    x0 = np.array([[0, np.pi, 0, 0]])
    ref = O.ddp_solve_batch("cartpole", p, O.ddp_config(max_iter=3, horizon_steps=N, with_input_constraint=1), x0,
                            np.zeros((1, N, 1)), u_lo=lo, u_hi=hi)
    solver = gpu.DDPSolver("cartpole", params=p, batch_capacity=1)
    c = solver.config()
    c.horizon_steps, c.max_iter, c.with_input_constraint = N, 3, True
    with pytest.raises(gpu.NmpcB200Error):  # limits not set yet
        solver.solve(0.0, x0[0], np.zeros((N, 1)))
    solver.setInputLimitsFunc(lambda t: (lo, hi))
    solver.solve(0.0, x0[0], np.zeros((N, 1)))
    np.testing.assert_allclose(solver.trace()[0, :, 1], ref["trace"][0, :, 1], rtol=1e-12)
    assert _rel_u(solver.controlData().u_list, ref["u"]).max() <= U_TOL_REF
    assert _rel_u(solver.k_list(), ref["k"]).max() <= 1e-7
    assert _rel_u(solver.K_list().reshape(1, N, -1), ref["K"]).max() <= 1e-7
    if N == 200:
        np.testing.assert_allclose(solver.trace()[0, :, 1],
                                   [991.8952423095, 882.7146741324, 850.4060100400, 843.1435488411], rtol=1e-11)


def test_constrained_batch_parity(gpu):
    """Control-limited DDP on a random batch, tighter limits so that many steps clamp."""
    B, N = 256, 100
    p = O.default_params("cartpole")
    lo, hi = np.array([-6.0]), np.array([9.0])
    x0 = O.cartpole_x0(B, 12)
    u_init = np.zeros((B, N, 1))
    ref = O.ddp_solve_batch("cartpole", p, O.ddp_config(max_iter=12, horizon_steps=N, with_input_constraint=1), x0,
                            u_init, u_lo=lo, u_hi=hi)
    solver = gpu.DDPSolver("cartpole", params=p, batch_capacity=B)
    c = solver.config()
    c.max_iter, c.with_input_constraint = 12, True
    solver.setInputLimitsFunc((lo, hi))
    solver.solve_batch(0.0, x0, u_init)
    assert np.array_equal(solver.iterations(), ref["iters"])
    assert np.array_equal(solver.status(), ref["status"])
    assert np.array_equal(solver.n_forward(), ref["n_fwd"])
    K = solver.K_list().reshape(B, N, -1)
    assert (K == 0).any() and (ref["K"] == 0).any(), "test input does not clamp"
    assert np.array_equal(K == 0, ref["K"] == 0), "clamped sets differ"
    assert _rel_u(solver.controlData().u_list, ref["u"]).max() <= U_TOL_REF
    assert np.max(np.abs(solver.cost() - ref["cost"]) / np.abs(ref["cost"])) <= COST_TOL_REF


@pytest.mark.parametrize("mode", ["ref", "fixed"])
def test_full_size_batches_are_independent_of_the_kernel_variant(gpu, mode):
    """BASELINE.json configs[1] (B = 4096) and a config-5 sweep point (B = 65536, which takes the large-batch line
    search kernel and many CTAs per SM), through size-independent properties:
      * instances never interact, so the first 512 instances solved ALONE (small-batch kernel variants, verified
        against the oracle above) must give bit-identical trajectories, costs and counters inside the big batches;
      * a sample of the big batch against the oracle;
      * every accepted step lowers the cost (trace), iteration counters within bounds."""
    kw = dict(max_iter=10)
    if mode == "fixed":
        kw.update(k_rel_norm_thre=0.0, cost_update_thre=0.0)
    N, sub = 100, 512
    p = O.default_params("cartpole")
    small = gpu.DDPSolver("cartpole", params=p, batch_capacity=sub)
    for k, v in kw.items():
        setattr(small.config(), k, v)
    x0_all = O.cartpole_x0(65536, 65536)
    small.solve_batch(0.0, x0_all[:sub], np.zeros((sub, N, 1)))
    want_u, want_cost = small.controlData().u_list, small.cost()
    want_it, want_fwd = small.iterations(), small.n_forward()
    ref = O.ddp_solve_batch("cartpole", p, O.ddp_config(horizon_steps=N, **kw), x0_all[:64], np.zeros((64, N, 1)))
    for B in (4096, 65536):
        big = gpu.DDPSolver("cartpole", params=p, batch_capacity=B)
        for k, v in kw.items():
            setattr(big.config(), k, v)
        big.solve_batch(0.0, x0_all[:B], np.zeros((B, N, 1)))
        u = big.controlData().u_list
        if B <= 4096:
            # same kernel variants as the small batch (lanes K2, split K3): bit-identical
            np.testing.assert_array_equal(u[:sub], want_u)
            np.testing.assert_array_equal(big.cost()[:sub], want_cost)
            np.testing.assert_array_equal(big.iterations()[:sub], want_it)
            np.testing.assert_array_equal(big.n_forward()[:sub], want_fwd)
        else:
            # the large-batch variants (thread-per-instance K2: (Fx^T Vxx) Fx instead of Fx^T (Vxx Fx)) differ at
            # rounding level
            assert _rel_u(u[:sub], want_u).max() <= (U_TOL_REF if mode == "ref" else U_TOL_FIXED)
            np.testing.assert_allclose(big.cost()[:sub], want_cost, rtol=COST_TOL_REF if mode == "ref" else COST_TOL_FIXED)
            np.testing.assert_array_equal(big.iterations()[:sub], want_it)
            if mode == "ref":
                np.testing.assert_array_equal(big.n_forward()[:sub], want_fwd)
        assert _rel_u(u[:64], ref["u"]).max() <= (U_TOL_REF if mode == "ref" else U_TOL_FIXED)
        assert np.array_equal(big.iterations()[:64], ref["iters"])
        it = big.iterations()
        assert it.min() >= 1 and it.max() <= 10 and (mode == "ref" or np.all(it == 10))
        tr = big.trace()  # [B, 11, 9]: rows of accepted steps carry a lower cost than the row before
        cost_rows = tr[:, :, 1]
        for r in range(1, 11):
            accepted = (tr[:, r, 4] > 0) & (tr[:, r, 6] > 0) & (r <= it)  # alpha set and cost_update_actual > 0
            assert np.all(cost_rows[accepted, r] < cost_rows[accepted, r - 1])
        assert np.isfinite(big.cost()).all()
        big.close()


@pytest.mark.parametrize("env", [{"NMPC_B200_BWD_LANES": "0"}, {"NMPC_B200_BWD_LANES": "2"}, {"NMPC_B200_BWD_LANES_TPC": "2"},
                                 {"NMPC_B200_BWD_FUSED": "0"}, {"NMPC_B200_BWD_QUAD": "1"}, {"NMPC_B200_BWD_GS": "4"},
                                 {"NMPC_B200_FWD_SPLIT": "0"}, {"NMPC_B200_FWD_SPLIT": "0", "NMPC_B200_FWD_GA": "1"},
                                 {"NMPC_B200_FWD_SPLIT": "0", "NMPC_B200_FWD_GA": "16"}, {"NMPC_B200_TILE": "1"}])
def test_every_kernel_variant_agrees_with_the_default(gpu, env, monkeypatch):
    """The engine's other kernel variants behind their environment switches against the default (K1+K2 with four lanes
    per instance, K3 split over rollout / cost / loader warps), with and without input limits.  Same arithmetic =>
    bit-identical controls, costs and counters: the shuffle exchange and two tiles per CTA of the lanes K2, every K3
    variant, the persistent tile kernel.  The thread-per-instance K2 family (fused, three-kernel pipeline, column
    split over 4 warps, in-warp cooperative) associates (Fx^T Vxx) Fx instead of Fx^T (Vxx Fx) and is held to the M-ref
    tolerances."""
    p = O.default_params("cartpole")
    B, Nh = 200, 100
    x0, u0 = O.cartpole_x0(B, 21), np.zeros((B, Nh, 1))

    def run(box):
        s = gpu.DDPSolver("cartpole", params=p, batch_capacity=B)
        c = s.config()
        c.max_iter, c.with_input_constraint = 8, box
        s.setInputLimitsFunc((np.array([-15.0]), np.array([15.0])))
        s.solve_batch(0.0, x0, u0)
        out = (s.controlData().u_list, s.cost(), s.iterations(), s.n_forward(), s.n_backward(), s.status())
        s.close()
        return out

    base = [run(False), run(True)]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    exact = not any(k in env for k in ("NMPC_B200_BWD_GS", "NMPC_B200_BWD_FUSED", "NMPC_B200_BWD_QUAD")) and env.get(
        "NMPC_B200_BWD_LANES") != "0"
    for want, box in zip(base, (False, True)):
        got = run(box)
        if exact:
            for a, b in zip(got, want):
                np.testing.assert_array_equal(a, b)
        else:
            assert _rel_u(got[0], want[0]).max() <= U_TOL_REF
            np.testing.assert_allclose(got[1], want[1], rtol=COST_TOL_REF)
            for a, b in zip(got[2:], want[2:]):
                np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("N", [1, 2, 3, 7])
def test_tiny_horizons(gpu, N):
    """Horizons shorter than the kernels' shared-memory rings (producer / consumer ring of the fused K2, operand rings
    of K3, chunked K0), unconstrained and BoxQP-constrained, ragged batch."""
    p = O.default_params("cartpole")
    B = 5
    x0, u0 = O.cartpole_x0(B, 40 + N), np.zeros((B, N, 1))
    for box in (0, 1):
        ref = O.ddp_solve_batch("cartpole", p, O.ddp_config(horizon_steps=N, max_iter=6, with_input_constraint=box), x0, u0,
                                u_lo=np.array([-2.0]), u_hi=np.array([2.0]))
        s = gpu.DDPSolver("cartpole", params=p, batch_capacity=B)
        c = s.config()
        c.horizon_steps, c.max_iter, c.with_input_constraint = N, 6, bool(box)
        s.setInputLimitsFunc((np.array([-2.0]), np.array([2.0])))
        s.solve_batch(0.0, x0, u0)
        assert np.array_equal(s.iterations(), ref["iters"]) and np.array_equal(s.status(), ref["status"])
        assert _rel_u(s.controlData().u_list, ref["u"]).max() <= U_TOL_REF
        assert np.max(np.abs(s.cost() - ref["cost"]) / np.abs(ref["cost"])) <= 1e-11
        s.close()


def test_input_limits_that_change_along_the_horizon(gpu):
    """setInputLimitsFunc with a genuine function of time (DDPSolver.h:282-285; evaluated at every t_i by backwardPass,
    DDPSolver.hpp:470): the limits tighten from +-15 N to +-3 N over the horizon.  Against the oracle driven with the
    same table; the constant-limit call must keep working after it; the device MPC loop runs with such limits too."""
    p = O.default_params("cartpole")
    B, N, t0 = 48, 100, 0.3
    x0, u0 = O.cartpole_x0(B, 77), np.zeros((B, N, 1))
    dt = p[0]

    def limits(t):
        w = 15.0 - 12.0 * min(max((t - t0) / (N * dt), 0.0), 1.0)
        return np.array([-w]), np.array([0.5 * w])

    lo = np.array([limits(t0 + i * dt)[0] for i in range(N)])
    hi = np.array([limits(t0 + i * dt)[1] for i in range(N)])
    cfg = O.ddp_config(horizon_steps=N, max_iter=8, with_input_constraint=1)
    ref = O.ddp_solve_cartpole_tv_limits(p, cfg, x0, u0, lo, hi, t0=t0)

    s = gpu.DDPSolver("cartpole", params=p, batch_capacity=B)
    c = s.config()
    c.max_iter, c.with_input_constraint = 8, True
    s.setInputLimitsFunc(limits)
    s.solve_batch(t0, x0, u0)
    assert np.array_equal(s.iterations(), ref["iters"]) and np.array_equal(s.status(), ref["status"])
    assert _rel_u(s.controlData().u_list, ref["u"]).max() <= U_TOL_REF
    assert np.max(np.abs(s.cost() - ref["cost"]) / np.abs(ref["cost"])) <= COST_TOL_REF
    k = s.k_list()[:, :, 0]  # the feedforward term respects each step's own box (lo_i - u_i <= k_i <= hi_i - u_i)
    # the device MPC loop takes such limits as per-tick tables (tests/test_mpc_gpu.py covers its results)
    log = s.run_mpc(t0, x0, u0, n_ticks=2, tick_dt=dt)
    assert np.isfinite(log["u"]).all()
    # constant limits afterwards: same result as a fresh solver
    s.setInputLimitsFunc((np.array([-15.0]), np.array([15.0])))
    s.solve_batch(0.0, x0, u0)
    ref_c = O.ddp_solve_batch("cartpole", p, O.ddp_config(horizon_steps=N, max_iter=8, with_input_constraint=1), x0, u0,
                              u_lo=np.array([-15.0]), u_hi=np.array([15.0]))
    assert _rel_u(s.controlData().u_list, ref_c["u"]).max() <= U_TOL_REF
    del k


def test_two_devices_in_one_process(gpu):
    """Handles on different devices are independent (c_api.h): the kernels that need more than 48 KB of dynamic shared
    memory (FMPC sweeps, column-split K2) must have their function attribute set on EVERY device.  Skipped on a
    single-GPU box."""
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    from test_quadrotor_gpu import N as NQ, hover_inputs, quadrotor_x0

    x0 = O.cartpole_x0(64, 1)
    outs = []
    for dev in (0, 1):
        s = gpu.FmpcSolver("cartpole", batch_capacity=64, device=dev)
        s.config().max_iter = 3
        v = s.make_variable(64)
        v.reset(0.0, 0.0, 0.0, 1.0, 1.0)
        s.solve_batch(0.0, x0, v)
        outs.append(s.variable().u_list.copy())
        q = gpu.DDPSolver("quadrotor", batch_capacity=64, device=dev)
        q.config().horizon_steps, q.config().max_iter = NQ, 2
        q.solve_batch(0.0, quadrotor_x0(64, 1), hover_inputs(64))
        outs.append(q.cost().copy())
    np.testing.assert_array_equal(outs[0], outs[2])
    np.testing.assert_array_equal(outs[1], outs[3])
