"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol of
include/nmpc_b200/c_api.h, and refuses to work without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nmpc_b200", "c_api.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nmpc_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(nmpc):
    from nmpc_b200 import _capi

    declared = _declared_symbols()
    assert len(declared) >= 25
    L = nmpc.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in c_api.h but not exported by libnmpc_b200.so"
    assert sorted(_capi.EXPORTED_SYMBOLS) == declared


def test_version_and_registry(nmpc):
    assert nmpc.lib().nmpc_b200_version() == 100
    names = nmpc.model_names()
    assert "cartpole" in names
    assert nmpc.model_dims("cartpole") == (4, 1, 4, 14)
    p = nmpc.model_default_params("cartpole")
    np.testing.assert_allclose(p, [0.01, 1.0, 0.5, 2.0, 0.1, 1.0, 0.01, 0.1, 0.01, 0.1, 1.0, 0.01, 0.1, 0.0])
    with pytest.raises(nmpc.NmpcB200Error) as e:
        nmpc.model_dims("no_such_problem")
    assert e.value.code == 3


def test_config_defaults_match_reference(nmpc):
    """DDPSolver.h:47-110 and FmpcSolver.h:58-89 defaults through the C ABI and the Python mirror."""
    from nmpc_b200 import _capi

    s = _capi.DdpConfigStruct()
    nmpc.lib().nmpc_b200_ddp_config_default(C.byref(s))
    d = nmpc.DDPConfiguration()
    assert (s.horizon_steps, s.max_iter, s.reg_type, s.with_input_constraint, s.n_alpha) == (100, 500, 1, 0, 11)
    assert (s.initial_lambda, s.initial_dlambda, s.lambda_factor, s.lambda_min, s.lambda_max) == (1e-4, 1.0, 1.6, 1e-6,
                                                                                                 1e10)
    assert (s.k_rel_norm_thre, s.lambda_thre, s.cost_update_ratio_thre, s.cost_update_thre) == (1e-4, 1e-5, 0.0, 1e-7)
    np.testing.assert_allclose(list(s.alpha_list)[:11], 10.0 ** np.linspace(0, -3, 11), rtol=1e-15)
    assert bytes(d.to_struct()) == bytes(s)
    f = _capi.FmpcConfigStruct()
    nmpc.lib().nmpc_b200_fmpc_config_default(C.byref(f))
    assert (f.horizon_steps, f.max_iter, f.check_nan, f.update_barrier_eps, f.enable_line_search) == (100, 10, 1, 1, 0)
    assert f.kkt_error_thre == 1e-4
    assert bytes(nmpc.FmpcConfiguration().to_struct()) == bytes(f)


def test_no_cpu_fallback(nmpc):
    """Without a CUDA device the product path fails loudly instead of computing on the CPU."""
    if nmpc.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(nmpc.NmpcB200Error) as e:
        nmpc.DDPSolver("cartpole", batch_capacity=4)
    assert e.value.code == 4
    assert "no CPU fallback" in e.value.message
    with pytest.raises(nmpc.NmpcB200Error):
        nmpc.model_eval("cartpole", 0.0, np.zeros((1, 4)), np.zeros((1, 1)))


def test_product_does_not_touch_the_oracle():
    """Nothing under nmpc_b200/ or include/ may import, include or link the oracle."""
    bad = []
    for base in ("nmpc_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "_obj" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if f.endswith((".so", ".o", ".log", ".pyc")):
                    continue
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"oracle_lib|liboracle|oracle/|#include\s*[\"<].*oracle", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
