"""The C++ host mirror (include/mlb200.hpp): compiles against the C ABI everywhere; on a GPU box
the reference's own assertions (restated in tests/cpp/test_hostapi.cpp) must hold."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_hostapi")


def build_exe():
    from madronalib_b200 import api, build
    if not os.path.exists(api.LIB_PATH):
        build.build()
    libdir = os.path.dirname(api.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_hostapi.cpp"), "-o", EXE,
           "-L", libdir, "-lmlb200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_cpp_host_header_compiles_and_refuses_cpu_fallback():
    from madronalib_b200 import api
    build_exe()
    if api.device_count() > 0:
        pytest.skip("GPU visible: covered by the gpu test")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 77, r.stdout + r.stderr
    assert "failed loudly" in r.stdout


@pytest.mark.gpu
def test_cpp_reference_assertions_on_gpu(gpu):
    build_exe()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASSED" in r.stdout
