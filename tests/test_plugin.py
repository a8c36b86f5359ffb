"""A user-supplied problem functor drops in from OUTSIDE the library (north_star; the reference binds any
std::shared_ptr<DDPProblem> at run time, DDPSolver.h:255): tests/plugin/pendulum_plugin.cu is compiled by nvcc into its
own shared library with the four lines include/nmpc_b200/plugin.h documents, loaded next to libnmpc_b200.so with
nmpc_b200_load_plugin, and solved through the C++ facade against the oracle instantiated on the same functor."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "tests", "plugin", "libpendulum_plugin.so")
BIN = os.path.join(ROOT, "tests", "cpp", "test_plugin")


def _newer(target, *sources):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in sources)


@pytest.fixture(scope="module")
def plugin_so(nmpc):
    """Step 3 of include/nmpc_b200/plugin.h (cross-compiles without a GPU)."""
    src = os.path.join(ROOT, "tests", "plugin", "pendulum_plugin.cu")
    deps = [src, os.path.join(ROOT, "tests", "plugin", "pendulum.h"), os.path.join(ROOT, "nmpc_b200", "libnmpc_b200.so")]
    if not _newer(PLUGIN, *deps):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
                        "--expt-relaxed-constexpr", "-ccbin", "g++", "-Xcompiler", "-fPIC", "-shared",
                        "-I" + os.path.join(ROOT, "include"), src, "-o", PLUGIN, "-L" + os.path.join(ROOT, "nmpc_b200"),
                        "-lnmpc_b200", "-Xlinker", "-rpath," + os.path.join(ROOT, "nmpc_b200")], check=True)
    return PLUGIN


@pytest.fixture(scope="module")
def plugin_bin(nmpc):
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "test_plugin.cpp"), "-o", BIN, "-L" + os.path.join(ROOT, "nmpc_b200"),
                    "-lnmpc_b200", "-Wl,-rpath," + os.path.join(ROOT, "nmpc_b200")], check=True)
    return BIN


def _parse(out):
    d = {}
    for line in out.splitlines():
        k, _, v = line.partition(" ")
        d[k] = v
    return d


def test_plugin_builds_loads_and_registers(plugin_so, plugin_bin, nmpc):
    """No GPU needed: the registrar of the plugin runs at load time and the functor is known to the library."""
    r = subprocess.run([plugin_bin, plugin_so, "--no-solve"], capture_output=True, text=True)
    d = _parse(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert d["known_before"] == "0" and d["load_rc"] == "0" and d["load_again"] == "0"
    assert int(d["load_missing"]) == 2  # NMPC_B200_ERR_RUNTIME with the loader's message
    assert d["known_after"] == "1 dims 2 1 0 9"
    # the Python mirror sees it too
    assert "pendulum" not in nmpc.model_names()
    nmpc.load_plugin(plugin_so)
    assert "pendulum" in nmpc.model_names() and nmpc.model_dims("pendulum") == (2, 1, 0, 9)
    if nmpc.device_count() == 0:
        with pytest.raises(nmpc.NmpcB200Error):  # no CPU fallback for plugin functors either
            nmpc.DDPSolver("pendulum", batch_capacity=4)


@pytest.mark.gpu
def test_plugin_functor_against_oracle(plugin_so, plugin_bin, gpu):
    r = subprocess.run([plugin_bin, plugin_so], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    d = _parse(r.stdout)
    assert float(d["rel_du"]) <= 1e-9 and float(d["rel_dcost"]) <= 1e-12, d
    assert d["iters_equal"] == "1" and d["status_equal"] == "1" and int(d["n_converged"]) > 0
