"""BoxQP<4> inside the control-limited DDP backward pass (DDPSolver.hpp:450-497), CPU side: the oracle against the
REFERENCE's own DDPSolver<12, 4> + BoxQP.h on the quadrotor functor (tests/golden/reference_ddp_quadrotor.npz), and the
reference algorithm's sensitivity to the last bit of its input, which bounds what tests/test_quadrotor_gpu.py can ask of
the CUDA path after the second iteration."""
import numpy as np
import pytest

from test_quadrotor_gpu import GOLDEN_BOX, _half_ulp, _oracle_box, _rel_u


@pytest.mark.parametrize("max_iter", [2, 6])
def test_oracle_boxqp_nu4_is_the_reference_headers_bit_for_bit(max_iter):
    """BoxQP<4> in the limited backward pass: the oracle against the REFERENCE's DDPSolver<12, 4> run on the same functor
    (tests/golden/make_golden_ddp_quadrotor.py).  Identical to the last bit, six iterations deep, all 64 instances."""
    g = GOLDEN_BOX
    r = _oracle_box(max_iter, g["x0"])
    np.testing.assert_array_equal(r["u"], g[f"it{max_iter}/u"])
    np.testing.assert_array_equal(r["cost_list"], g[f"it{max_iter}/cost_list"])
    np.testing.assert_array_equal(r["trace"], g[f"it{max_iter}/trace"])
    assert np.array_equal(r["n_trace"], g[f"it{max_iter}/n_trace"])


def test_reference_boxqp_nu4_splits_on_the_last_bit():
    """Why no implementation with another rounding can follow the reference past its second iteration here: the
    reference algorithm, solved again from x0 changed by half an ulp, keeps every instance for two iterations (1e-10) and
    by the sixth has moved a quarter or more of them by O(0.01 .. 0.1) in u (free-set choice by `x == lower`,
    BoxQP.h:189-191, on inputs that forwardPass left within an ulp of a limit)."""
    g = GOLDEN_BOX
    x1 = _half_ulp(g["x0"], 1)
    assert _rel_u(_oracle_box(2, x1)["u"], g["it2/u"]).max() < 1e-10
    moved = _rel_u(_oracle_box(6, x1)["u"], g["it6/u"])
    assert (moved > 1e-6).sum() >= 16 and moved.max() > 1e-2, ((moved > 1e-6).sum(), moved.max())
