"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol that
include/mlb200.h declares, agrees with the Python op table, validates graphs, designs
coefficients exactly like the reference, and refuses to compute without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

from madronalib_b200 import api, graph
from madronalib_b200.graph import OP_TABLE, GraphSpec, Layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(api.LIB_PATH):
        from madronalib_b200 import build
        build.build()
    return api.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mlb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mlb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(L):
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"libmlb200.so does not export {s}"
    assert L.mlb_abi_version() == 2


def test_op_table_matches_library(L):
    for name, (op, nin, nst, nco) in OP_TABLE.items():
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert L.mlb_op_info(op, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)) == 0
        assert (a.value, b.value, c.value) == (nin, nst, nco), name
        assert L.mlb_op_name(op).decode() == name
    assert L.mlb_op_info(9999, None, None, None) != 0


def test_graph_layout_and_validation(L):
    g = graph.graph_fm3_fdn8()
    lay = Layout()
    n = g.n_nodes
    so, co = (ctypes.c_int32 * n)(), (ctypes.c_int32 * n)()
    assert L.mlb_graph_layout(g.c_nodes(), n, ctypes.byref(lay), so, co) == 0
    assert (lay.n_state_words, lay.n_coef_words, lay.n_inputs) == (g.n_state, g.n_coef, 1)
    assert list(so) == g.offsets()[0] and list(co) == g.offsets()[1]
    # forward reference is rejected
    bad = g.c_nodes()
    bad[6].inp[0] = 9
    assert L.mlb_graph_layout(bad, n, ctypes.byref(lay), None, None) == 1
    assert b"earlier node" in L.mlb_last_error()
    # unknown op
    bad = g.c_nodes()
    bad[2].op = 777
    assert L.mlb_graph_layout(bad, n, ctypes.byref(lay), None, None) == 1


def test_validation_of_the_paired_and_feedback_nodes(L):
    lay = Layout()
    # FEEDBACK_WRITE must name an earlier FEEDBACK_READ
    g = graph.GraphSpec()
    x = g.input(0)
    fb = g.feedback_read()
    g.feedback_write(fb, g.node("ADD", x, fb))
    assert L.mlb_graph_layout(g.c_nodes(), g.n_nodes, ctypes.byref(lay), None, None) == 0
    bad = g.c_nodes()
    bad[g.n_nodes - 1].iarg = 0  # the INPUT node
    assert L.mlb_graph_layout(bad, g.n_nodes, ctypes.byref(lay), None, None) == 1
    assert b"FEEDBACK_READ" in L.mlb_last_error()
    # HALFBAND_UP_2 must read a HALFBAND_UP
    h = graph.GraphSpec()
    up = h.node("HALFBAND_UP", h.input(0))
    h.output(h.node("HALFBAND_UP_2", up))
    assert L.mlb_graph_layout(h.c_nodes(), h.n_nodes, ctypes.byref(lay), None, None) == 0
    assert lay.n_state_words == 9
    bad = h.c_nodes()
    bad[2].inp[0] = 0
    assert L.mlb_graph_layout(bad, h.n_nodes, ctypes.byref(lay), None, None) == 1


def test_again_nodes_share_the_words_of_their_functor(L):
    """MLB_AGAIN: a further call of an earlier functor owns no words; the C layout, the Python layout and both oracles
    agree; the rules of mlb200.h are enforced."""
    from madronalib_b200 import workloads as wl
    g = wl.functor_case("upsample2x_osc", 4).spec
    n = g.n_nodes
    lay = Layout()
    so, co = (ctypes.c_int32 * n)(), (ctypes.c_int32 * n)()
    assert L.mlb_graph_layout(g.c_nodes(), n, ctypes.byref(lay), so, co) == 0
    st, cf, ns, nc = g.offsets()
    assert list(so) == st and list(co) == cf and (lay.n_state_words, lay.n_coef_words) == (ns, nc) == (21, 4)
    again = [i for i in range(n) if g.again_target(i) >= 0]
    assert len(again) == 2 and all(so[i] == so[g.again_target(i)] and co[i] == co[g.again_target(i)] for i in again)

    def rejects(edit, words):
        bad = g.c_nodes()
        edit(bad)
        assert L.mlb_graph_layout(bad, n, ctypes.byref(lay), None, None) == 1
        assert words in L.mlb_last_error(), L.mlb_last_error()

    a = again[0]
    rejects(lambda b: setattr(b[a], "iarg", -1 - a), b"earlier node")        # itself / a later node
    rejects(lambda b: setattr(b[a], "iarg", -1 - 0), b"same op")             # node 0 is the INPUT
    rejects(lambda b: setattr(b[again[1]], "iarg", -1 - again[0]), b"same op")
    # an AGAIN node cannot be the target of another one
    h = graph.GraphSpec()
    x = h.input(0)
    s1 = h.node("SINE", x)
    s2 = h.again(s1, x)
    with pytest.raises(ValueError):
        h.again(s2, x)
    bad = graph.GraphSpec()
    bad.ops, bad.ins, bad.iargs = list(h.ops) + [h.ops[s1]], list(h.ins) + [h.ins[s1]], list(h.iargs) + [-1 - s2]
    assert L.mlb_graph_layout(bad.c_nodes(), 4, ctypes.byref(lay), None, None) == 1
    # functors with a ring in delay memory, stateless ops and the paired / feedback nodes cannot be called again
    for name, nin in (("INTEGER_DELAY", 1), ("ALLPASS_PB", 2), ("ADD", 2), ("DOWN2X_IN", 1), ("FDN8", 1)):
        q = graph.GraphSpec()
        ins = [q.input(k) for k in range(nin)]
        first = q.node(name, *ins)
        q.ops.append(q.ops[first]), q.ins.append(q.ins[first]), q.iargs.append(-1 - first)
        assert L.mlb_graph_layout(q.c_nodes(), q.n_nodes, ctypes.byref(lay), None, None) == 1, name
        assert b"cannot be called again" in L.mlb_last_error()
    # ... while the half-band filters can (the stages of an Upsampler / Downsampler run theirs several times per vector)
    q = graph.GraphSpec()
    x = q.input(0)
    u1 = q.node("HALFBAND_UP", x)
    u1b = q.node("HALFBAND_UP_2", u1)
    u2 = q.again(u1, u1b)
    q.output(q.node("HALFBAND_UP_2", u2))
    assert L.mlb_graph_layout(q.c_nodes(), q.n_nodes, ctypes.byref(lay), None, None) == 0 and lay.n_state_words == 9
    # ... and so can a functor with a member row but no ring (LinearGlide keeps mCurrVec)
    q = graph.GraphSpec()
    first = q.node("GLIDE", q.input(0))
    q.output(q.again(first, q.input(1)))
    assert L.mlb_graph_layout(q.c_nodes(), q.n_nodes, ctypes.byref(lay), None, None) == 0
    assert lay.n_state_words == graph.OP_INFO[graph.OP_ID["GLIDE"]][1]


def test_no_gpu_means_loud_failure_not_fallback(L):
    if api.device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(api.MlbError) as e:
        api.VoiceGraph(graph.graph_sine_lopass_gain(), 8)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)
    with pytest.raises(api.MlbError):
        api.map_host("sin", np.zeros((1, 64), np.float32))
    with pytest.raises(api.MlbError) as e:  # the Voice bank too
        api.VoiceBank(48000.0, [1], [0.0], [0.0], [7.0])
    assert e.value.code == 3


def test_coefficient_design_matches_reference_golden(L):
    gold = np.load(os.path.join(GOLD, "coeffs.npz"))
    om = gold["coef_omega"]
    for kind, extra in (("lopass", (0.5,)), ("hipass", (0.5,)), ("bandpass", (0.3,)),
                        ("loshelf", (0.7, 1.4)), ("hishelf", (0.7, 1.4)), ("bell", (0.5, 1.41)),
                        ("onepole", ())):
        got = np.stack([api.coeffs(kind, float(o), *extra) for o in om])
        assert np.array_equal(got.view(np.uint32), gold["coef_" + kind].view(np.uint32)), kind
    assert np.float32(api.db_to_gain(6.0)) == gold["db_to_gain_6"]
    assert np.float32(api.coeffs_dcblocker(0.045)) == gold["dcblocker_0045"]
    # SURVEY 8c spot values
    assert [float(x).hex() for x in api.coeffs("lopass", 0.1, 0.5)] == \
        ["0x1.0663920000000p-2", "-0x1.b0e64e0000000p-3", "0x1.55057a0000000p-4"]
    assert [float(x).hex() for x in api.coeffs("onepole", 0.01)] == \
        ["0x1.f2e1c00000000p-5", "0x1.e0d1e40000000p-1"]


def test_coefficient_design_matches_port(L, port):
    for kind, args in (("lopass", (0.123, 0.4)), ("bell", (0.2, 0.3, 2.0)), ("onepole", (0.3,))):
        assert np.array_equal(api.coeffs(kind, *args).view(np.uint32),
                              port.coeffs(kind, *args).view(np.uint32))


def test_fdn_coefs(L):
    c = api.coeffs_fdn8([67, 73, 91, 103, 127, 151, 173, 263], [0.1] * 8, [0.5] * 8)
    # FDN::setDelaysInSamples: len = max(1, int(time - 64))   (MLDSPFilters.h:1173-1183)
    assert list(c[24:]) == [3, 9, 27, 39, 63, 87, 109, 199]
    assert np.array_equal(c[:8], np.repeat(api.coeffs("onepole", 0.1)[0], 8))


def test_graphspec_builders():
    g = graph.graph_chain256(256)
    assert g.n_nodes == 5 + 1 + 256 and g.n_in == 0 and g.n_out == 1
    assert graph.graph_sine_lopass_gain().n_state == 3
    with pytest.raises(ValueError):
        GraphSpec().node("LOPASS", 3)
