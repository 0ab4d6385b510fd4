"""Edge cases of the GPU path: error behaviour of the C ABI, denormals, extreme inputs,
fast mode tolerance, long launches, every generator / one-pole family node through both kernels."""
import numpy as np
import pytest

from madronalib_b200 import workloads as wl
from madronalib_b200.graph import GraphSpec, OP_ID
from tests.common import assert_same_bits, assert_state_equal, run_gpu

pytestmark = pytest.mark.gpu


def test_error_codes(gpu):
    w = wl.config_a(64)
    g = gpu.VoiceGraph(w.spec, 64)
    with pytest.raises(gpu.MlbError) as e:  # INPUT graph without input
        gpu._check(gpu.lib().mlb_graph_process_host(g._h, None, None, None, 2))
    assert e.value.code == 1
    with pytest.raises(gpu.MlbError):
        gpu._check(gpu.lib().mlb_graph_process_host(g._h, None, None, None, 0))
    g.close()
    with pytest.raises(gpu.MlbError):  # zero voices
        gpu.VoiceGraph(w.spec, 0)
    bad = GraphSpec()
    bad.ops, bad.ins, bad.iargs, bad.outs = [OP_ID["LOPASS"]], [(0, -1, -1)], [0], [0]  # self reference
    with pytest.raises(gpu.MlbError):
        gpu.VoiceGraph(bad, 8)
    with pytest.raises(gpu.MlbError):  # stateful op through the map entry point
        gpu.map_host("lopass", np.zeros((1, 64), np.float32))
    # FDN8 needs its coefficients (delay lengths) before it can allocate rings
    f = wl.config_4(8)
    gf = gpu.VoiceGraph(f.spec, 8)
    with pytest.raises(gpu.MlbError):
        gf.process_host(f.inputs(1), 1)
    gf.close()
    # delay nodes: the maxDelay coefficient sizes the ring and must be sane
    d = wl.functor_case("allpass_frac", 8)
    gd = gpu.VoiceGraph(d.spec, 8)
    c = d.coef.copy()
    c[2, 3] = np.float32(-5.0)
    with pytest.raises(gpu.MlbError) as e:
        gd.set_coefs(c)
    assert e.value.code == 1 and "maxDelay" in str(e.value)
    c[2, 3] = np.float32(np.nan)
    with pytest.raises(gpu.MlbError):
        gd.set_coefs(c)
    gd.set_coefs(d.coef)  # a good upload still works afterwards
    gd.process_host(d.inputs(2), 2)
    gd.close()
    # Voice bank argument checks
    vb = gpu.VoiceBank(48000.0, [1, 2], [0.0, 0.0], [0.0, 0.0], [7.0, 7.0])
    with pytest.raises(gpu.MlbError):
        gpu._check(gpu.lib().mlb_voices_process_host(vb._h, None, None, 4, 0xFF))
    ev = np.zeros((2, 2), wl.VOICE_EVENTS_DTYPE)
    with pytest.raises(gpu.MlbError):
        gpu._check(gpu.lib().mlb_voices_process_host(vb._h, ev.ctypes.data, None, 2, 0xFF))
    assert gpu.lib().mlb_voices_process_host(vb._h, ev.ctypes.data, None, 2, 0) == 0  # no rows wanted: fine
    vb.close()
    with pytest.raises(gpu.MlbError):
        gpu.VoiceBank(0.0, [1], [0.0], [0.0], [7.0])  # sample rate must be positive


def test_denormals_are_honoured(gpu, port):
    """The reference runs with IEEE denormals unless UsingFlushDenormalsToZero is active
    (MLDSPUtils.h:51-96): a decaying OnePole must pass through the denormal range identically."""
    g = GraphSpec()
    g.output(g.node("ONEPOLE", g.input(0)))
    V, T = 40, 12
    coef = g.new_coefs(V)
    coef[0], coef[1] = 1e-3, 0.5  # y = 1e-3 x + 0.5 y  -> halves every sample once x = 0
    st = g.new_state(V)
    st[0] = np.float32(1e-30).view(np.uint32)
    inp = np.zeros((T, 1, V, 64), np.float32)
    inp[0, 0, :, 0] = 1e-38
    w = wl.Workload("denorm", g, V, coef, st)
    po, _, ps = port.run(g, V, T, inp, st, coef)
    go, _, gs, _ = run_gpu(gpu, w, T, inp)
    tiny = np.finfo(np.float32).tiny
    assert ((np.abs(po) > 0) & (np.abs(po) < tiny)).any(), "test must exercise denormals"
    assert_same_bits(go, po)
    assert_state_equal(gs, ps)


@pytest.mark.parametrize("flags", [0, 2])
def test_extreme_frequencies(gpu, port, flags):
    """cvtps2dq overflow, NaN, infinities and negative frequencies in the phase accumulator."""
    w = wl.config_a(64)
    T = 3
    inp = w.inputs(T)
    bad = np.array([0.5, 0.75, -0.5, 1e9, -1e9, np.inf, -np.inf, np.nan, 2.0 ** -40, -0.0], np.float32)
    inp[1, 0, :, 10:20] = bad
    po, _, ps = port.run(w.spec, 64, T, inp, w.state, w.coef)
    go, _, gs, _ = run_gpu(gpu, w, T, inp, flags=flags)
    assert_same_bits(go, po)
    assert_state_equal(gs, ps)


def test_fast_mode_stated_tolerance(gpu, port):
    """MLB_GRAPH_FAST allows FMA contraction: not bit-exact, relative error small and bounded."""
    V, T = 128, 64
    w = wl.config_a(V)
    inp = w.inputs(T)
    po, _, _ = port.run(w.spec, V, T, inp, w.state, w.coef, nthreads=4)
    go, _, _, name = run_gpu(gpu, w, T, inp, flags=gpu.FLAG_FAST)
    assert name.endswith("(fast)")
    err = np.abs(go - po).max() / np.abs(po).max()
    assert err < 1e-4, err          # stated tolerance of fast mode on config A over 64 blocks
    assert not np.array_equal(go, po)  # it really is a different rounding sequence


def test_long_launch_many_blocks(gpu, port):
    V, T = 70, 300
    w = wl.config_3(V)
    inp = w.inputs(T)
    po, _, ps = port.run(w.spec, V, T, inp, w.state, w.coef, nthreads=4)
    go, _, gs, _ = run_gpu(gpu, w, T, inp)
    assert_same_bits(go, po)
    assert_state_equal(gs, ps)


@pytest.mark.parametrize("name", ["SAW", "PHASOR", "TICK", "NOISE"])
def test_generators_fused_or_generic(gpu, port, name):
    g = GraphSpec()
    g.output(g.node(name) if name == "NOISE" else g.node(name, g.input(0)))
    V, T = 45, 4
    st = g.new_state(V)
    st[0] = np.arange(V, dtype=np.uint32) * 977 if name != "TICK" else np.linspace(0, 1, V, dtype=np.float32).view(np.uint32)
    w = wl.Workload(name, g, V, g.new_coefs(V), st)
    inp = None if name == "NOISE" else wl.freq_rows(V, T)
    po, _, ps = port.run(g, V, T, inp, st, w.coef)
    go, _, gs, _ = run_gpu(gpu, w, T, inp)
    assert_same_bits(go, po, name)
    assert_state_equal(gs, ps, name)


def test_pulse_and_onepole_family_in_one_graph(gpu, port):
    """PulseGen(freq, width) -> DCBlocker -> Differentiator -> Integrator, plus a second output."""
    g = GraphSpec()
    f, wdt = g.input(0), g.input(1)
    p = g.node("PULSE", f, wdt)
    d = g.node("DCBLOCKER", p)
    df = g.node("DIFFERENTIATOR", d)
    y = g.node("INTEGRATOR", df)
    g.output(y, d)
    V, T = 33, 5
    coef = g.new_coefs(V)
    coef[g.coef_slot(d)] = gpu.coeffs_dcblocker(0.045)
    coef[g.coef_slot(y)] = 0.001
    inp = np.concatenate([wl.freq_rows(V, T), np.full((T, 1, V, 64), 0.3, np.float32)], axis=1)
    st = g.new_state(V)
    w = wl.Workload("pulse", g, V, coef, st)
    po, pm, ps = port.run(g, V, T, inp, st, coef, want_mix=True, mix_mode=1)
    go, gm, gs, name = run_gpu(gpu, w, T, inp, want_mix=True, splits=(2, 3))
    assert name.startswith("generic")
    assert_same_bits(go, po)
    assert_same_bits(gm, pm)
    assert_state_equal(gs, ps)


def test_state_roundtrip_migrates_a_voice_mid_stream(gpu, port):
    """get_state / set_state use the reference's member layout: run 3 blocks on the GPU, move the
    state to the CPU oracle for 2 blocks, move it back, and compare with an uninterrupted run."""
    V = 50
    w = wl.config_3(V)
    inp = w.inputs(8)
    full, _, _ = port.run(w.spec, V, 8, inp, w.state, w.coef)
    g = gpu.VoiceGraph(w.spec, V)
    g.set_coefs(w.coef)
    g.set_state(w.state)
    a, _ = g.process_host(np.ascontiguousarray(inp[0:3]), 3)
    st = g.get_state()
    b, _, st2 = port.run(w.spec, V, 2, np.ascontiguousarray(inp[3:5]), st, w.coef)
    g.set_state(st2)
    c, _ = g.process_host(np.ascontiguousarray(inp[5:8]), 3)
    g.close()
    assert_same_bits(np.concatenate([a, b, c]), full)
