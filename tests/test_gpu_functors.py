"""GPU parity for SURVEY 8(f) row 2 -- the rest of the L2 functor set and the Aaltoverb example
chain, through the C ABI, against the C port (which the CPU suite pins to the compiled reference)."""
import numpy as np
import pytest

from madronalib_b200 import workloads as wl
from tests.common import assert_same_bits, assert_state_equal, run_gpu

pytestmark = pytest.mark.gpu

# Peak / RMS end in sqrtApprox = x * rsqrt(x): a 12-bit CPU-defined approximation on the reference
# side (MLDSPMathSSE.h:84-85).  Their OUTPUT is compared with this relative tolerance; their state
# (the exact recurrence) is compared bit for bit.
APPROX_RTOL = 4e-4


@pytest.mark.parametrize("name", wl.FUNCTOR_CASES)
@pytest.mark.parametrize("n_voices", [40, 97])
def test_functor_bit_exact(gpu, port, name, n_voices):
    w = wl.functor_case(name, n_voices)
    T = 24
    inp = w.inputs(T)
    po, _, ps = port.run(w.spec, n_voices, T, inp, w.state, w.coef)
    go, _, gs, kname = run_gpu(gpu, w, T, inp, splits=(5, 7, 12))
    assert kname.startswith("generic"), kname
    if name in ("peak", "rms"):
        np.testing.assert_allclose(go, po, rtol=APPROX_RTOL, atol=1e-12)
    else:
        assert_same_bits(go, po, name)
    assert_state_equal(gs, ps, name)


def test_functor_against_reference_itself(gpu, ref):
    for name in ("adsr", "allpass_pb", "glide", "fractional_delay_var"):
        w = wl.functor_case(name, 33)
        inp = w.inputs(10)
        ro, _, rs = ref.run(w.spec, 33, 10, inp, w.state, w.coef)
        go, _, gs, _ = run_gpu(gpu, w, 10, inp)
        assert_same_bits(go, ro, name)
        assert_state_equal(gs, rs, name)


@pytest.mark.parametrize("name,dmax", [("integer_delay_var", 2200.0), ("fractional_delay_var", 2100.0),
                                       ("pitchbend_delay", 1500.0)])
def test_delay_beyond_the_ring(gpu, port, name, dmax):
    """Delay times past the ring (and NaN / negative ones) make the reference's per-sample loop read
    slots of the current block before it writes them; the block-wise kernel must return the same."""
    w = wl.functor_case(name, 40)
    T = 12
    inp = w.inputs(T)
    d = inp[:, 1]
    d *= np.float32(dmax / max(float(d.max()), 1.0))       # sweep well past maxDelay (ring 1024 / 2048)
    d[3, 5, 10:20] = -7.5
    d[4, 6, 0:64] = np.nan
    d[5, 7, 16] = 1e12
    po, _, ps = port.run(w.spec, 40, T, inp, w.state, w.coef)
    go, _, gs, _ = run_gpu(gpu, w, T, inp, splits=(4, 8))
    assert_same_bits(go, po, name)
    assert_state_equal(gs, ps, name)


@pytest.mark.parametrize("n_voices,n_blocks", [(6, 130), (70, 40)])
def test_aaltoverb_bit_exact(gpu, port, n_voices, n_blocks):
    """examples/audio-and-midi/reverb.cpp as a voice graph (10 Allpass<PitchbendableDelay>, two
    PitchbendableDelay feedback lines, two LinearGlides, one-block feedback edges)."""
    w = wl.config_6(n_voices)
    inp = w.inputs(n_blocks)
    po, pm, ps = port.run(w.spec, n_voices, n_blocks, inp, w.state, w.coef, want_mix=True, mix_mode=1)
    go, gm, gs, kname = run_gpu(gpu, w, n_blocks, inp, want_mix=True, splits=(n_blocks // 3, n_blocks - n_blocks // 3))
    assert kname.startswith("generic")
    assert_same_bits(go, po, "aaltoverb out")
    assert_same_bits(gm, pm, "aaltoverb mix")
    assert_state_equal(gs, ps, "aaltoverb state")
    assert np.sqrt((go[-5:] ** 2).mean()) > 1e-3


def test_clear_delays_restarts_the_tail(gpu, port):
    w = wl.functor_case("allpass_frac", 40)
    inp = w.inputs(8)
    g = gpu.VoiceGraph(w.spec, 40)
    try:
        g.set_coefs(w.coef)
        g.set_state(w.state)
        a, _ = g.process_host(inp, 8)
        g.clear_delays()
        g.set_state(w.state)
        b, _ = g.process_host(inp, 8)
        assert g.delay_bytes > 0
    finally:
        g.close()
    assert_same_bits(a, b, "after clear_delays")


@pytest.mark.parametrize("name", wl.FUNCTOR_CASES + ("aaltoverb",))
def test_functor_matches_committed_reference_golden(gpu, name):
    """tests/golden/functors.npz was written by the compiled reference (make_golden.py)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "functors.npz"))
    inp = gold[name + "_in"]
    T, V = inp.shape[0], inp.shape[2]
    w = wl.config_6(V) if name == "aaltoverb" else wl.functor_case(name, V)
    w.coef, w.state = gold[name + "_coef"], gold[name + "_state0"]
    go, _, gs, _ = run_gpu(gpu, w, T, inp)
    if name in ("peak", "rms"):
        np.testing.assert_allclose(go, gold[name + "_out"], rtol=APPROX_RTOL, atol=1e-12)
    else:
        assert_same_bits(go, gold[name + "_out"], name)
    assert_state_equal(gs, gold[name + "_state1"], name)


@pytest.mark.parametrize("stages", [1, 2, 5, 64])
def test_stage_pipeline_gives_identical_bits(gpu, port, monkeypatch, stages):
    """The interpreter cuts the program into stages run by different CTAs (rows cross through channel
    planes, feedback rows through delay memory); any cut must give the same bits."""
    monkeypatch.setenv("MLB_STAGES", str(stages))
    for w, T in ((wl.config_6(70), 30), (wl.config_5(40, 64), 6), (wl.functor_case("feedback", 50), 12)):
        inp = w.inputs(T)
        po, pm, ps = port.run(w.spec, w.n_voices, T, inp, w.state, w.coef, want_mix=True, mix_mode=1)
        go, gm, gs, kname = run_gpu(gpu, w, T, inp, want_mix=True, flags=gpu.FLAG_FORCE_GENERIC,
                                    splits=(T // 2, T - T // 2))
        assert "stages" in kname, kname
        assert_same_bits(go, po, w.name)
        assert_same_bits(gm, pm, w.name + " mix")
        assert_state_equal(gs, ps, w.name)


@pytest.mark.parametrize("seed", range(8))
def test_random_graphs_bit_exact(gpu, port, monkeypatch, seed):
    """Random DAGs over most of the op table through the interpreter (default stage count, then 3 stages),
    against the port -- which the CPU suite pins to the compiled reference on the same graphs.

    Seeds are limited to graphs in which no op observes the SIGN BIT of a NaN: x86 produces the negative
    default NaN (0xFFC00000), sm_100 the positive canonical one (0x7FFFFFFF), so sign(NaN), signBit(NaN) or a
    bitwise select on a NaN legitimately differ (seed 8: feedback row = -inf, NaN -> SIGN)."""
    w = wl.random_graph_workload(seed, 37 + seed, 26, hw_approx=False)
    T = 9
    inp = w.inputs(T)
    po, pm, ps = port.run(w.spec, w.n_voices, T, inp, w.state, w.coef, want_mix=True, mix_mode=1)
    for stages in (None, 3):
        if stages:
            monkeypatch.setenv("MLB_STAGES", str(stages))
        go, gm, gs, kname = run_gpu(gpu, w, T, inp, want_mix=True, splits=(4, 5))
        assert kname.startswith("generic"), kname
        assert_same_bits(go, po, "random graph %d (%s)" % (seed, kname))
        assert_state_equal(gs, ps, "random graph %d (%s)" % (seed, kname))
        finite = np.isfinite(pm) & np.isfinite(gm)
        assert np.array_equal(gm[finite].view(np.uint32), pm[finite].view(np.uint32))

