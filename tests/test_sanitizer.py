"""compute-sanitizer over the hand-synchronised kernels (mbarrier rings between producer / loader / consumer warps,
cp.async + mbarrier.arrive.noinc, group barriers, the lambda-retry votes): memcheck, synccheck and racecheck must
report nothing on tools/sanitize_cases.py.  Logs of the round's last run are kept under profiles/r2_sanitizer/ (summary: profiles/r2_sanitizer.md)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


DIAG_LIB = os.path.join(ROOT, "nmpc_b200", "libnmpc_b200_diag.so")


def _sanitize(tool, cases, timeout, library=None):
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer is not installed")
    env = dict(os.environ)
    if library:
        env["NMPC_B200_LIBRARY"] = library
    r = subprocess.run([exe, "--tool", tool, sys.executable, os.path.join(ROOT, "tools", "sanitize_cases.py"), *cases],
                       capture_output=True, text=True, timeout=timeout, env=env)
    out = r.stdout + r.stderr
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = "diag_" if library else ""
    with open(os.path.join(ROOT, "gpurun_out", f"sanitizer_{tool}_{tag}{'_'.join(cases)}.txt"), "w") as f:
        f.write(out)
    assert r.returncode == 0, out[-3000:]
    for c in cases:
        assert f"case {c} done" in out, out[-3000:]
    return out


def _run(tool, cases, timeout):
    out = _sanitize(tool, cases, timeout)
    assert "ERROR SUMMARY: 0 errors" in out, out[-3000:]


@pytest.mark.parametrize("cases", [["lanes", "retry", "tiny"], ["variants"], ["fmpc"], ["quad", "wide"]])
def test_memcheck(gpu, cases):
    _run("memcheck", cases, 900)


@pytest.mark.parametrize("cases", [["lanes", "retry", "tiny"], ["variants"], ["fmpc"], ["quad", "wide"]])
def test_synccheck(gpu, cases):
    _run("synccheck", cases, 900)


@pytest.mark.parametrize("cases", [["lanes", "retry", "tiny"], ["variants"], ["fmpc"]])
def test_racecheck(gpu, cases):
    """racecheck follows mbarrier arrive / wait, the group barriers and the CTA barriers of these kernels, but not the
    arrive the HARDWARE performs when a loader lane's cp.async copies complete (cp.async.mbarrier.arrive.noinc): for the
    rings filled that way it intermittently reports the asynchronous write against the consumer's read.  So:
      * in the product build every reported hazard must be one of those (a cp.async write on one side);
      * the diagnostic build, whose loaders wait for their copies and arrive themselves (the ONLY difference,
        NMPC_B200_LOADER_WAITS in ddp_kernels.cuh; 3 % slower, bit-identical results), must be clean."""
    out = _sanitize("racecheck", cases, 1500)
    records = out.split("Race reported between")[1:]
    other = [r[:400] for r in records if "cpAsync" not in r.split("=========", 3)[0] + r.split("=========", 3)[1]]
    assert not other, other[:3]
    if not os.path.exists(DIAG_LIB):
        pytest.skip("diagnostic library not built (make -C nmpc_b200/csrc diag)")
    out = _sanitize("racecheck", cases, 1500, library=DIAG_LIB)
    assert "RACECHECK SUMMARY: 0 hazards displayed (0 errors, 0 warnings)" in out, out[-3000:]
