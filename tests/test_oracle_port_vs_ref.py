"""The plain-C port against the compiled reference on fresh random inputs (only where
oracle/_ref exists: the authoring container, or a GPU box that received the built .so)."""
import numpy as np
import pytest

from madronalib_b200 import workloads as wl
from madronalib_b200.graph import OP_TABLE, GraphSpec
from tests.common import assert_state_equal


def test_this_is_the_reference(ref):
    # SURVEY appendix A sizes: DSPVector, Lopass, OnePole, SineGen, IntegerDelay, FDN<8>
    assert [ref.sizeof(i) for i in range(6)] == [256, 20, 12, 4, 48, 2560]


@pytest.mark.parametrize("seed", [0, 1])
def test_all_stateless_ops(ref, port, seed):
    rng = np.random.default_rng(seed)
    V, T = 5, 2
    for name, (_, nin, nst, nco) in OP_TABLE.items():
        if not (30 <= OP_TABLE[name][0] < 100):  # MLB_OP_MAP_FIRST .. MLB_OP_MAP_END
            continue
        g = GraphSpec()
        g.output(g.node(name, *[g.input(k) for k in range(nin)]))
        x = (rng.standard_normal((T, nin, V, 64)) * 10.0 ** rng.integers(-3, 4)).astype(np.float32)
        a, _, _ = ref.run(g, V, T, x, g.new_state(V), g.new_coefs(V))
        b, _, _ = port.run(g, V, T, x, g.new_state(V), g.new_coefs(V))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), name


@pytest.mark.parametrize("make,T", [
    (lambda: wl.config_a(70), 5), (lambda: wl.config_2("bell", 40), 4),
    (lambda: wl.config_2("hishelf", 40), 4), (lambda: wl.config_3(50), 4),
    (lambda: wl.config_4(24), 20), (lambda: wl.config_5(6, 256), 3)])
def test_configs_random_state(ref, port, make, T):
    w = make()
    rng = np.random.default_rng(3)
    st = w.state.copy()
    # start mid-stream: random filter state (phases stay as set)
    for i, op in enumerate(w.spec.ops):
        from madronalib_b200.graph import OP_INFO, OP_NAME
        if OP_NAME[op] in ("LOPASS", "HIPASS", "BANDPASS", "LOSHELF", "HISHELF", "BELL", "ONEPOLE"):
            s0 = w.spec.state_slot(i)
            n = OP_INFO[op][1]
            st[s0:s0 + n] = (rng.standard_normal((n, w.n_voices)) * 0.1).astype(np.float32).view(np.uint32)
    inp = w.inputs(T)
    a, am, ast = ref.run(w.spec, w.n_voices, T, inp, st, w.coef, want_mix=True)
    b, bm, bst = port.run(w.spec, w.n_voices, T, inp, st, w.coef, want_mix=True, nthreads=3)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(am.view(np.uint32), bm.view(np.uint32))
    assert_state_equal(ast, bst)


def test_reference_chain_loop_equals_graph(ref):
    """The CPU-baseline loop (struct Voice{SineGen; Lopass}) computes exactly the graph result."""
    w = wl.config_a(64)
    T = 4
    inp = w.inputs(T)
    want, _, _ = ref.run(w.spec, 64, T, inp, w.state, w.coef)
    phase = w.state[0].copy()
    ic = w.state[1:3].view(np.float32).copy()
    got, sec = ref.chain_sine_lopass_gain(np.ascontiguousarray(inp[:, 0]), np.ascontiguousarray(w.coef[0:3]),
                                          np.ascontiguousarray(w.coef[3]), phase, ic, 2)
    assert np.array_equal(got.view(np.uint32), want[:, 0].view(np.uint32)) and sec > 0


@pytest.mark.parametrize("name", wl.FUNCTOR_CASES)
def test_functor_cases(ref, port, name):
    """every added functor, port == compiled reference, bit for bit, with the
    port run split over three calls (state and delay memory carried inside the oracle)."""
    w = wl.functor_case(name, 40)
    T = 24
    inp = w.inputs(T)
    a, _, ast = ref.run(w.spec, w.n_voices, T, inp, w.state, w.coef)
    b, _, bst = port.run(w.spec, w.n_voices, T, inp, w.state, w.coef, splits=(5, 7, 12), nthreads=2)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert_state_equal(ast, bst)
    assert np.abs(a).max() > 0


def test_aaltoverb_graph_is_the_example(ref, port):
    """graph_aaltoverb() == the body of examples/audio-and-midi/reverb.cpp driven directly
    (mlref_aaltoverb), and the port agrees with both."""
    w = wl.config_6(6)
    T = 120
    inp = w.inputs(T)
    a, _, ast = ref.run(w.spec, w.n_voices, T, inp, w.state, w.coef)
    b, _, bst = port.run(w.spec, w.n_voices, T, inp, w.state, w.coef, splits=(50, 70))
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert_state_equal(ast, bst)
    for v in (0, 3, 5):
        size2 = float(w.coef[w.spec.coef_slot(2), v])
        fb = float(w.coef[w.spec.coef_slot(3), v])
        o, _ = ref.aaltoverb(inp[:, :, v, :], size2, fb, 4800.0)
        assert np.array_equal(o.view(np.uint32), a[:, :, v, :].view(np.uint32))
    assert np.sqrt((a[-10:] ** 2).mean()) > 1e-3  # the tail is ringing, not silent


def test_upsample2x_graph_is_the_higher_order_function(ref):
    """HALFBAND_DOWN(fn(HALFBAND_UP(x)), fn(HALFBAND_UP_2(x))) == the reference's Upsample2xFunction<1>
    (MLDSPFunctional.h:114-160) called directly with fn(v) = clamp(v * drive, -1, 1)."""
    w = wl.functor_case("upsample2x_clip", 12)
    T = 20
    inp = w.inputs(T)
    a, _, _ = ref.run(w.spec, 12, T, inp, w.state, w.coef)
    for v in (0, 7, 11):
        o = ref.upsample2x_clip(inp[:, 0, v, :], float(w.coef[0, v]))
        assert np.array_equal(o.view(np.uint32), a[:, 0, v, :].view(np.uint32))


def test_downsample2x_graph_is_the_higher_order_function(ref):
    """DOWN2X_OUT(fn(DOWN2X_IN(x))) == the reference's Downsample2xFunction<1> (MLDSPFunctional.h:166-223)
    called directly with fn(v) = clamp(v * drive, -1, 1)."""
    w = wl.functor_case("downsample2x_clip", 12)
    T = 21
    inp = w.inputs(T)
    a, _, _ = ref.run(w.spec, 12, T, inp, w.state, w.coef)
    for v in (0, 7, 11):
        o = ref.downsample2x_clip(inp[:, 0, v, :], float(w.coef[0, v]))
        assert np.array_equal(o.view(np.uint32), a[:, 0, v, :].view(np.uint32))


@pytest.mark.parametrize("seed", range(16))
def test_random_graphs(ref, port, seed):
    """Random DAGs over most of the op table (generators, filters, delay functors, feedback edges, elementwise
    ops): the C port and the compiled reference agree bit for bit, NaNs included, with the port run split."""
    w = wl.random_graph_workload(seed, 21, 26)
    T = 9
    inp = w.inputs(T)
    a, _, ast = ref.run(w.spec, w.n_voices, T, inp, w.state, w.coef)
    b, _, bst = port.run(w.spec, w.n_voices, T, inp, w.state, w.coef, splits=(4, 5))
    from tests.common import assert_same_bits
    assert_same_bits(b, a, "random graph %d" % seed)
    assert_state_equal(bst, ast, "random graph %d" % seed)
