"""Host-side ring buffer + batched SignalProcessBuffer shim: CPU-only C++ test,
driven in lockstep with the reference's own DSPBuffer where /root/reference exists."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_ringbuffer_and_batched_process_buffer():
    exe = os.path.join(ROOT, "tests", "cpp", "test_ringbuffer")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_ringbuffer.cpp"), "-o", exe]
    have_ref = os.path.isdir(os.path.join(REF, "source", "DSP"))
    if have_ref:
        cmd += ["-DHAVE_REFERENCE", "-fno-strict-aliasing", "-include", "cstdint", "-include", "cstddef",
                "-I", os.path.join(REF, "source", "DSP"), "-I", os.path.join(REF, "source", "app"),
                "-I", os.path.join(REF, "include"), "-I", os.path.join(REF, "external")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    if have_ref:
        assert "checked against the reference" in r.stdout
