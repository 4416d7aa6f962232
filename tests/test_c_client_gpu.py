"""The C ABI used from plain C (no Python, no torch): compile tests/c_abi/kat_test.c with gcc,
link libgsfield.so, run the reference's known-answer vectors.  This is the call pattern of the
Rust FFI shim."""
import os
import subprocess

import pytest

from conftest import ROOT

LIBDIR = os.path.join(ROOT, "gstools-core_b200", "gstools_core")
SRC = os.path.join(ROOT, "tests", "c_abi", "kat_test.c")


def _build(tmp_path):
    exe = str(tmp_path / "kat_test")
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    SRC, "-L", LIBDIR, "-lgsfield", "-lm", "-Wl,-rpath," + LIBDIR, "-o", exe], check=True)
    return exe


def test_c_client_compiles_and_links(tmp_path):
    # header is valid C11, every used symbol resolves (runs on the CPU box: no device => exit code 3)
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode in (0, 3), r.stdout + r.stderr


@pytest.mark.gpu
def test_c_client_kat(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "C ABI KAT OK" in r.stdout
