"""The header-only C++ host API (include/icebin/regrid_b200.hpp) compiles with plain g++ against the C ABI (CPU),
and behaves like the reference's C++ API on a GPU (gpu)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_api")


def build_exe():
    from icebin_b200 import build
    build.build()
    src = os.path.join(ROOT, "tests", "cpp", "test_api.cpp")
    so = os.path.join(ROOT, "icebin_b200", "libicebin_b200.so")
    hdrs = [os.path.join(ROOT, "include", "icebin_b200.h")] + [os.path.join(ROOT, "include", "icebin", f)
                                                             for f in os.listdir(os.path.join(ROOT, "include", "icebin"))]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(f) for f in [src, so] + hdrs):
        subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L", os.path.join(ROOT, "icebin_b200"), "-licebin_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "icebin_b200"), "-Wl,-rpath,/usr/local/cuda/lib64"])
    return EXE


def test_cpp_header_compiles_and_links():
    out = subprocess.check_output([build_exe(), "--compile-only"], text=True)
    assert "linked ok" in out


@pytest.mark.gpu
def test_cpp_api_on_gpu():
    r = subprocess.run([build_exe()], capture_output=True, text=True)
    assert r.returncode == 0 and "C++ API ok" in r.stdout, r.stdout + r.stderr
