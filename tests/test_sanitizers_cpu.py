"""Host runtime under AddressSanitizer + UndefinedBehaviorSanitizer.

`make -C gstools-core_b200/csrc asan` builds libgsfield_asan.so (host code instrumented; ~2 min, so
it is not part of build()).  The test then re-runs the CPU-only host-logic and ABI suites in a child
interpreter with that library (GSF_LIB) and the sanitizer runtimes preloaded; any report fails it.
Skipped when the instrumented library has not been built (set GSF_BUILD_ASAN=1 to build it here).
On a GPU box the same recipe runs the device tests: see profiles/sanitizer_r2.md."""
import glob
import os
import subprocess
import sys

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "gstools-core_b200", "csrc")
ASAN_LIB = os.path.join(ROOT, "gstools-core_b200", "gstools_core", "libgsfield_asan.so")


def _runtime(name):
    hits = sorted(glob.glob("/usr/lib/x86_64-linux-gnu/%s.so.*" % name) + glob.glob("/usr/lib64/%s.so.*" % name))
    return hits[0] if hits else None


def test_host_logic_under_asan_ubsan():
    sources = glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(CSRC, "*.inc")) + glob.glob(os.path.join(CSRC, "*.cpp")) + [os.path.join(ROOT, "include", "gsfield.h")]
    stale = os.path.exists(ASAN_LIB) and os.path.getmtime(ASAN_LIB) < max(os.path.getmtime(f) for f in sources)
    if not os.path.exists(ASAN_LIB) or stale:
        if os.environ.get("GSF_BUILD_ASAN") != "1":
            pytest.skip("libgsfield_asan.so %s (make -C gstools-core_b200/csrc asan, or GSF_BUILD_ASAN=1)"
                        % ("is older than the sources" if stale else "not built"))
        subprocess.run(["make", "-C", CSRC, "asan"], check=True)
    asan, ubsan = _runtime("libasan"), _runtime("libubsan")
    if not asan or not ubsan:
        pytest.skip("sanitizer runtimes not installed")
    env = dict(os.environ, GSF_LIB=ASAN_LIB, LD_PRELOAD="%s %s" % (asan, ubsan),
               ASAN_OPTIONS="detect_leaks=0:protect_shadow_gap=0", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "not gpu",
                        os.path.join(ROOT, "tests", "test_host_logic_cpu.py"), os.path.join(ROOT, "tests", "test_abi_cpu.py")],
                       capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    report = r.stdout + r.stderr
    assert r.returncode == 0, report[-3000:]
    assert "ERROR: AddressSanitizer" not in report and "runtime error:" not in report, report[-3000:]
