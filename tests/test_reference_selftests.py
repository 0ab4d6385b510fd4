"""The reference's OWN DSP tests (Tests/dspOpsTest.cpp, dspGensTest.cpp, dspFiltersTest.cpp, dspBufferTest.cpp),
compiled in place with exactly the flag set the oracle is built with (oracle/Makefile): the oracle's build of the
reference is one the reference's own assertions accept.  CPU-only, needs /root/reference (skipped elsewhere)."""
import os
import subprocess
import tempfile

import pytest

REF = "/root/reference"


def test_reference_dsp_tests_pass_under_the_oracle_flags():
    if not os.path.isdir(os.path.join(REF, "Tests")):
        pytest.skip("no /root/reference here")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "reftests")
        cmd = ["g++", "-std=c++17", "-O2", "-fno-strict-aliasing", "-ffp-contract=off", "-include", "cstdint",
               "-include", "cstddef"]
        for d in ("include", "source/DSP", "source/app", "external/ffft", "external", "Tests"):
            cmd += ["-I", os.path.join(REF, d)]
        cmd += [os.path.join(REF, "Tests", f) for f in ("tests.cpp", "dspOpsTest.cpp", "dspGensTest.cpp",
                                                        "dspFiltersTest.cpp", "dspBufferTest.cpp")]
        cmd += ["-lpthread", "-o", exe]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "All tests passed" in r.stdout, r.stdout[-2000:]
