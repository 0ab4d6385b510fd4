"""Contract E, events -> signals -> chain in one call (mlb_synth_process_host): only the 72-byte event
records go up; the Voice bank (MLEventsToSignals.cpp:383-470) writes the rows the graph reads straight
into the graph's input buffer on the device, and the graph's rows / mix bus come back.  Checked against
port(bank) -> port(graph), the composition a reference synth performs per vector (MLSynth.h:36-60)."""
import os

import numpy as np
import pytest

from madronalib_b200 import workloads as wl
from tests.common import assert_same_bits, assert_state_equal

pytestmark = pytest.mark.gpu

NTHREADS = max(1, len(os.sched_getaffinity(0)))


@pytest.fixture(scope="module")
def port_bank():
    from oracle import bindings
    return bindings.port_voice_bank()


def _expect(port, port_bank, w, ev, prm, planes, want_mix=True):
    T, V = ev.shape
    rows, _ = port_bank.run(48000.0, *prm, ev, nthreads=NTHREADS)
    inp = np.ascontiguousarray(rows[:, planes])
    out, _, st = port.run(w.spec, V, T, inp, w.state, w.coef, nthreads=NTHREADS)
    mix = None
    if want_mix:
        _, mix, _ = port.run(w.spec, V, T, inp, w.state, w.coef, want_out=False, want_mix=True, mix_mode=1,
                             nthreads=NTHREADS)
    return out, mix, st


@pytest.mark.parametrize("chunk,calls", [(8, (20,)), (0, (20,)), (3, (7, 13))])
def test_events_to_fused_chain(gpu, port, port_bank, monkeypatch, chunk, calls):
    """Config A reading its frequency row from the bank's kPitch row; time chunks of 8 / one chunk / two
    successive calls with ragged chunks (bank and chain state carry across chunks and calls)."""
    if chunk:
        monkeypatch.setenv("MLB_SYNTH_CHUNK_BLOCKS", str(chunk))
    else:
        monkeypatch.delenv("MLB_SYNTH_CHUNK_BLOCKS", raising=False)
    V, T = 300, sum(calls)
    w = wl.config_a(V)
    ev = wl.synth_events(V, T, seed=21, density=0.3, ctl=0.2)
    prm = wl.synth_bank_params(V)
    want_out, want_mix, want_st = _expect(port, port_bank, w, ev, prm, [0])
    vb = gpu.VoiceBank(48000.0, *prm)
    g = gpu.VoiceGraph(w.spec, V)
    try:
        assert g.kernel_name.startswith("fused:")
        g.set_coefs(w.coef)
        g.set_state(w.state)
        outs, mixes, t0 = [], [], 0
        for n in calls:
            o, m = g.process_events_host(vb, np.ascontiguousarray(ev[t0:t0 + n]), want_out=True, want_mix=True)
            outs.append(o), mixes.append(m)
            t0 += n
            if chunk:
                assert g.last_host_slices == (n + chunk - 1) // chunk
        st = g.get_state()
    finally:
        g.close()
        vb.close()
    assert_same_bits(np.concatenate(outs), want_out, "events -> chain A rows")
    assert_same_bits(np.concatenate(mixes), want_mix, "events -> chain A mix bus")
    assert_state_equal(st, want_st, "events -> chain A state")
    assert np.abs(want_out).max() > 0.01


def test_events_to_chain_mix_only(gpu, port, port_bank, monkeypatch):
    """No output rows leave the device (contract E in, contract M out)."""
    monkeypatch.setenv("MLB_SYNTH_CHUNK_BLOCKS", "5")
    V, T = 131, 12
    w = wl.config_a(V)
    ev = wl.synth_events(V, T, seed=22, density=0.3, ctl=0.2)
    prm = wl.synth_bank_params(V)
    _, want_mix, want_st = _expect(port, port_bank, w, ev, prm, [0])
    vb = gpu.VoiceBank(48000.0, *prm)
    g = gpu.VoiceGraph(w.spec, V)
    try:
        g.set_coefs(w.coef)
        g.set_state(w.state)
        o, m = g.process_events_host(vb, ev, want_out=False, want_mix=True)
        st = g.get_state()
    finally:
        g.close()
        vb.close()
    assert o is None
    assert_same_bits(m, want_mix, "events -> mix bus only")
    assert_state_equal(st, want_st, "state")


def test_events_to_interpreted_graph(gpu, port, port_bank, monkeypatch):
    """A graph reading two Voice rows (pitch x 2^-2 -> SineGen -> x gate) through the interpreter; the
    bank generates rows 0 and 1 only."""
    from madronalib_b200.graph import GraphSpec, SINE_ZERO_PHASE
    monkeypatch.setenv("MLB_SYNTH_CHUNK_BLOCKS", "6")
    V, T = 71, 16
    gs = GraphSpec()
    pitch, gate = gs.input(0), gs.input(1)
    k = gs.param()
    gs.output(gs.node("MULTIPLY", gs.node("SINE", gs.node("MULTIPLY", pitch, k)), gate))
    coef, state = gs.new_coefs(V), gs.new_state(V)
    coef[0] = np.float32(0.25)
    state[0] = SINE_ZERO_PHASE
    w = wl.Workload("pitch_sine_gate", gs, V, coef, state)
    ev = wl.synth_events(V, T, seed=23, density=0.3, ctl=0.2)
    prm = wl.synth_bank_params(V)
    want_out, _, want_st = _expect(port, port_bank, w, ev, prm, [0, 1], want_mix=False)
    vb = gpu.VoiceBank(48000.0, *prm)
    g = gpu.VoiceGraph(gs, V)
    try:
        g.set_coefs(coef)
        g.set_state(state)
        o, _ = g.process_events_host(vb, ev, want_out=True)
        st = g.get_state()
    finally:
        g.close()
        vb.close()
    assert_same_bits(o, want_out, "events -> interpreted graph")
    assert_state_equal(st, want_st, "state")
    assert np.abs(want_out).max() > 0.01


def test_synth_argument_errors(gpu):
    from madronalib_b200.graph import GraphSpec
    V = 64
    w = wl.config_a(V)
    prm = wl.synth_bank_params(V)
    vb = gpu.VoiceBank(48000.0, *prm)
    vb_small = gpu.VoiceBank(48000.0, *wl.voice_bank_params(32))
    g = gpu.VoiceGraph(w.spec, V)
    no_in = GraphSpec()
    no_in.output(no_in.node("NOISE"))
    g2 = gpu.VoiceGraph(no_in, V)
    far = GraphSpec()
    far.output(far.node("SINE", far.input(9)))
    g3 = gpu.VoiceGraph(far, V)
    ev = wl.synth_events(V, 2)
    try:
        g.set_coefs(w.coef)
        with pytest.raises(gpu.MlbError, match="voices"):
            g.process_events_host(vb_small, ev)
        with pytest.raises(gpu.MlbError, match="no Voice row"):
            g2.process_events_host(vb, ev)
        with pytest.raises(gpu.MlbError, match="not a Voice row"):
            g3.process_events_host(vb, ev)
        with pytest.raises(ValueError):
            g.process_events_host(vb, ev[:, :10])
    finally:
        for h in (g, g2, g3, vb, vb_small):
            h.close()


def test_events_to_chain_full_size(gpu, port, port_bank, monkeypatch):
    """65 536 voices x 64 vectors, the default time chunks (8 x 8 vectors): output rows on a stride of
    voices against port(bank) -> port(graph) (voices are independent), the mix bus against the rows."""
    monkeypatch.delenv("MLB_SYNTH_CHUNK_BLOCKS", raising=False)
    V, T = 65536, 64
    w = wl.config_a(V)
    ev = wl.synth_events(V, T)
    prm = wl.synth_bank_params(V)
    vb = gpu.VoiceBank(48000.0, *prm)
    g = gpu.VoiceGraph(w.spec, V)
    try:
        g.set_coefs(w.coef)
        g.set_state(w.state)
        o, m = g.process_events_host(vb, ev, want_out=True, want_mix=True)
        assert g.last_host_slices == 8
    finally:
        g.close()
        vb.close()
    sel = np.arange(0, V, 61)
    rows, _ = port_bank.run(48000.0, *(p[sel] for p in prm), np.ascontiguousarray(ev[:, sel]), nthreads=NTHREADS)
    want, _, _ = port.run(w.spec, len(sel), T, np.ascontiguousarray(rows[:, 0:1]),
                          np.ascontiguousarray(w.state[:, sel]), np.ascontiguousarray(w.coef[:, sel]),
                          nthreads=NTHREADS)
    assert_same_bits(o[:, :, sel], want, "contract E 65536 x 64 (1075-voice sample)")
    ref_mix = o.astype(np.float64).sum(axis=2)
    tol = V * np.finfo(np.float32).eps * np.abs(o).sum(axis=2).max()
    assert np.abs(m - ref_mix).max() <= tol
