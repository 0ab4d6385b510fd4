"""mlb::VoiceRouter (include/mlb200_events.hpp) + Voice bank == the complete reference EventsToSignals
on MIDI phrases (tests/cpp/test_router.cpp).  CPU: bank = the C port.  GPU: bank = the CUDA kernel."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_router")


def build_exe():
    from madronalib_b200 import api, build
    from oracle import bindings
    if not os.path.exists(api.LIB_PATH):
        build.build()
    if not os.path.exists(bindings.PORT_LIB):
        bindings.build("port")
    if not os.path.exists(bindings.E2S_LIB):
        if os.path.isdir("/root/reference/source/app"):
            bindings.build("ref")
        else:
            pytest.skip("oracle/_ref/libmle2s.so not built (no /root/reference here)")
    dirs = [os.path.dirname(bindings.E2S_LIB), os.path.dirname(bindings.PORT_LIB), os.path.dirname(api.LIB_PATH)]
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_router.cpp"), "-o", EXE, "-lmle2s", "-lmlport", "-lmlb200"]
    for d in dirs:
        cmd += ["-L", d, "-Wl,-rpath," + d]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_router_plus_port_bank_equals_reference_events_to_signals():
    build_exe()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_router_plus_gpu_bank_equals_reference_events_to_signals(gpu):
    build_exe()
    r = subprocess.run([EXE, "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout + r.stderr
