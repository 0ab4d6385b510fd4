"""Voice-graph description shared by the C-ABI binding and the test oracles.

Pure Python + ctypes; no GPU, no oracle imports.  Mirrors ``include/mlb200.h``
(the op table is parsed from that header so the two can never drift).

A graph is what a user of the reference writes inside a ``SignalProcessFn`` with
functors (reference: examples/audio-and-midi/sine.cpp:21-43): generators,
filters and elementwise ops wired in a fixed DAG, evaluated once per 64-sample
block for every voice of a ``Bank`` (reference: source/DSP/MLDSPFunctional.h:321-360).
"""
from __future__ import annotations

import ctypes
import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

BLOCK = 64  # kFloatsPerDSPVector, reference source/DSP/MLDSPMath.h:8-9

_HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "mlb200.h")


def _parse_op_table() -> Dict[str, Tuple[int, int, int, int]]:
    """name -> (id, n_in, n_state, n_coef), parsed from MLB_OP_TABLE in mlb200.h."""
    with open(_HEADER, "r") as f:
        text = f.read()
    start = text.index("#define MLB_OP_TABLE(X)")
    end = text.index("typedef enum mlb_op")
    table = {}
    for m in re.finditer(r"X\(\s*([A-Z0-9_]+)\s*,\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*\)",
                         text[start:end]):
        table[m.group(1)] = tuple(int(m.group(i)) for i in range(2, 6))
    return table


def _parse_mem_table() -> Dict[str, Tuple[int, int]]:
    """name -> (n_rows, n_rings) of delay memory, parsed from MLB_OP_MEM_TABLE."""
    with open(_HEADER, "r") as f:
        text = f.read()
    start = text.index("#define MLB_OP_MEM_TABLE(X)")
    end = text.index("typedef enum mlb_op")
    return {m.group(1): (int(m.group(2)), int(m.group(3)))
            for m in re.finditer(r"X\(\s*([A-Z0-9_]+)\s*,\s*(\d+)\s*,\s*(\d+)\s*\)", text[start:end])}


def _parse_define(name: str) -> int:
    with open(_HEADER, "r") as f:
        return int(re.search(r"#define\s+" + name + r"\s+(\d+)", f.read()).group(1))


MAX_INS = _parse_define("MLB_MAX_INS")
ABI_VERSION = _parse_define("MLB_ABI_VERSION")
OP_TABLE = _parse_op_table()
OP_MEM = _parse_mem_table()
OP_ID = {name: v[0] for name, v in OP_TABLE.items()}
OP_NAME = {v[0]: name for name, v in OP_TABLE.items()}
OP_INFO = {v[0]: v[1:] for v in OP_TABLE.values()}  # id -> (n_in, n_state, n_coef)


class Node(ctypes.Structure):
    """struct mlb_node"""
    _fields_ = [("op", ctypes.c_int32), ("inp", ctypes.c_int32 * MAX_INS), ("iarg", ctypes.c_int32)]


class Layout(ctypes.Structure):
    """struct mlb_layout"""
    _fields_ = [("n_state_words", ctypes.c_int32), ("n_coef_words", ctypes.c_int32),
                ("n_inputs", ctypes.c_int32)]


@dataclass
class GraphSpec:
    """A validated node list plus its struct-of-arrays layout."""
    ops: List[int] = field(default_factory=list)
    ins: List[Tuple[int, int, int]] = field(default_factory=list)
    iargs: List[int] = field(default_factory=list)
    outs: List[int] = field(default_factory=list)

    # ---- construction (functional style, like chaining reference functors) ----
    def _add(self, name: str, *inputs: int, iarg: int = 0) -> int:
        op = OP_ID[name]
        n_in = OP_INFO[op][0]
        if len(inputs) != n_in:
            raise ValueError(f"{name} takes {n_in} inputs, got {len(inputs)}")
        idx = len(self.ops)
        for i in inputs:
            if not (0 <= i < idx):
                raise ValueError(f"{name}: input {i} is not an earlier node")
        pad = tuple(inputs) + (-1,) * (MAX_INS - len(inputs))
        self.ops.append(op)
        self.ins.append(pad)  # type: ignore[arg-type]
        self.iargs.append(iarg)
        return idx

    def input(self, k: int = 0) -> int:
        return self._add("INPUT", iarg=k)

    def param(self) -> int:
        return self._add("PARAM")

    def node(self, name: str, *inputs: int) -> int:
        return self._add(name.upper(), *inputs)

    def again(self, target: int, *inputs: int) -> int:
        """The functor of node `target` called once more in the same vector, on other inputs (MLB_AGAIN in mlb200.h:
        what happens to the functors of a process function that Upsample2xFunction runs twice per vector)."""
        if not (0 <= target < len(self.ops)) or self.again_target(target) >= 0:
            raise ValueError("again: target must be an earlier node that is not itself an `again` node")
        return self._add(OP_NAME[self.ops[target]], *inputs, iarg=-1 - target)

    def again_target(self, node: int) -> int:
        """The node whose functor `node` calls again, or -1."""
        if OP_NAME[self.ops[node]] in ("INPUT", "PARAM", "FEEDBACK_WRITE") or self.iargs[node] >= 0:
            return -1
        return -1 - self.iargs[node]

    def feedback_read(self) -> int:
        """The row stored by `feedback_write` on the previous block (zeros at start)."""
        return self._add("FEEDBACK_READ")

    def feedback_write(self, reader: int, src: int) -> int:
        if self.ops[reader] != OP_ID["FEEDBACK_READ"]:
            raise ValueError("feedback_write: target is not a FEEDBACK_READ node")
        return self._add("FEEDBACK_WRITE", src, iarg=reader)

    def output(self, *nodes: int) -> "GraphSpec":
        self.outs.extend(nodes)
        return self

    # ---- layout ----
    @property
    def n_nodes(self) -> int:
        return len(self.ops)

    @property
    def n_out(self) -> int:
        return len(self.outs)

    @property
    def n_in(self) -> int:
        ks = [self.iargs[i] + 1 for i, op in enumerate(self.ops) if op == OP_ID["INPUT"]]
        return max(ks) if ks else 0

    def offsets(self) -> Tuple[List[int], List[int], int, int]:
        st, co, ns, nc = [], [], 0, 0
        for i, op in enumerate(self.ops):
            t = self.again_target(i)
            if t >= 0:  # a further call of node t's functor: its words, none of its own
                st.append(st[t])
                co.append(co[t])
                continue
            st.append(ns)
            co.append(nc)
            ns += OP_INFO[op][1]
            nc += OP_INFO[op][2]
        return st, co, ns, nc

    @property
    def n_state(self) -> int:
        return self.offsets()[2]

    @property
    def n_coef(self) -> int:
        return self.offsets()[3]

    def state_slot(self, node: int, k: int = 0) -> int:
        return self.offsets()[0][node] + k

    def coef_slot(self, node: int, k: int = 0) -> int:
        return self.offsets()[1][node] + k

    # ---- ctypes views ----
    def c_nodes(self):
        arr = (Node * max(1, self.n_nodes))()
        for i in range(self.n_nodes):
            arr[i].op = self.ops[i]
            for k in range(MAX_INS):  # shorter tuples (hand-built specs) are padded with -1
                arr[i].inp[k] = self.ins[i][k] if k < len(self.ins[i]) else -1
            arr[i].iarg = self.iargs[i]
        return arr

    def c_outs(self):
        return (ctypes.c_int32 * max(1, self.n_out))(*self.outs)

    def new_state(self, n_voices: int) -> np.ndarray:
        """State of freshly constructed functors (zeros, except the idle markers below)."""
        st = np.zeros((max(1, self.n_state), n_voices), dtype=np.uint32)[: self.n_state]
        off = self.offsets()[0]
        for i, op in enumerate(self.ops):
            name = OP_NAME[op]
            if name == "ADSR":            # segment{off}, MLDSPFilters.h:694
                st[off[i] + 7] = 4
            elif name == "GLIDE":         # mVectorsRemaining{-1}, MLDSPGens.h:440
                st[off[i] + 2] = 0xFFFFFFFF
            elif name == "SAMPLE_GLIDE":  # mSamplesRemaining{-1}, MLDSPGens.h:524
                st[off[i] + 3] = 0xFFFFFFFF
            elif name == "TEMPO_LOCK":    # _omega{-1.f}, MLDSPFilters.h:1481
                st[off[i]] = 0xBF800000
        return st

    def new_coefs(self, n_voices: int) -> np.ndarray:
        return np.zeros((max(1, self.n_coef), n_voices), dtype=np.float32)[: self.n_coef]


# ---- ready-made graphs for the BASELINE.json configurations ----

SINE_ZERO_PHASE = 0xC0000000  # SineGen::clear(), reference MLDSPGens.h:375,379


def graph_sine_lopass_gain() -> GraphSpec:
    """Config 1 / Config A: gain * Lopass(SineGen(freq)); nodes: 0 in, 1 sine, 2 lopass, 3 gain, 4 mul."""
    g = GraphSpec()
    f = g.input(0)
    s = g.node("SINE", f)
    lp = g.node("LOPASS", s)
    k = g.param()
    y = g.node("MULTIPLY", lp, k)
    return g.output(y)


def graph_sine_svf(kind: str) -> GraphSpec:
    """Config 2: SineGen -> one of the SVF family ("Biquad" stand-ins)."""
    g = GraphSpec()
    f = g.input(0)
    s = g.node("SINE", f)
    y = g.node(kind.upper(), s)
    return g.output(y)


def graph_phasor_lopass_onepole() -> GraphSpec:
    """Config 3: PhasorGen -> Lopass -> OnePole (4 state words per voice)."""
    g = GraphSpec()
    f = g.input(0)
    p = g.node("PHASOR", f)
    lp = g.node("LOPASS", p)
    y = g.node("ONEPOLE", lp)
    return g.output(y)


def graph_fm3_fdn8() -> GraphSpec:
    """Config 4: 3-op FM (two modulators + carrier) into FDN<8>, stereo out.

    carrier freq = f * ((1 + i1*mod1) + i2*mod2), all by multiply/add on
    cyclesPerSample rows.  PARAM nodes: r1, r2, i1, i2, one.
    """
    g = GraphSpec()
    f = g.input(0)
    r1, r2, i1, i2, one = g.param(), g.param(), g.param(), g.param(), g.param()
    m1 = g.node("SINE", g.node("MULTIPLY", f, r1))
    m2 = g.node("SINE", g.node("MULTIPLY", f, r2))
    a = g.node("ADD", one, g.node("MULTIPLY", m1, i1))
    b = g.node("ADD", a, g.node("MULTIPLY", m2, i2))
    car = g.node("SINE", g.node("MULTIPLY", f, b))
    fl = g.node("FDN8", car)
    fr = g.node("FDN8_R", fl)
    return g.output(fl, fr)


def graph_fdn(size: int) -> Tuple["GraphSpec", Dict[str, List[int]]]:
    """FDN<SIZE> for any SIZE, written out with the nodes it is made of (MLDSPFilters.h:1162-1239): SIZE IntegerDelays fed
    by the vectors kept from the previous call (one-block feedback edges), the stereo sums, the Householder matrix as
    "minus 2/SIZE times the sum", SIZE OnePoles, the feedback gains, plus the input.  (SIZE = 8 also has the fused FDN8
    node and kernel.)  Same operation order as the reference, so the same bits.  Input plane 0; outputs sumL, sumR.
    Returns the graph and the node indices a caller sets coefficients on: 'delay', 'filter', 'gain' (per line),
    'zero' and 'k' (the constants 0 and 2.0f / SIZE)."""
    g = GraphSpec()
    x = g.input(0)
    zero, k = g.param(), g.param()
    readers = [g.feedback_read() for _ in range(size)]
    d = [g.node("INTEGER_DELAY", r) for r in readers]           # mDelayInputVectors[n] = mDelays[n](mDelayInputVectors[n])
    sum_r, sum_l = zero, zero                                   # DSPVector sumR, sumL: zero-filled
    for n in range(size & ~1):
        if n & 1:
            sum_l = g.node("ADD", sum_l, d[n])
        else:
            sum_r = g.node("ADD", sum_r, d[n])
    total = zero
    for n in range(size):
        total = g.node("ADD", total, d[n])
    total = g.node("MULTIPLY", total, k)                        # sumOfDelays *= DSPVector(2.0f / SIZE)
    filt, gain = [], []
    for n in range(size):
        f = g.node("ONEPOLE", g.node("SUBTRACT", d[n], total))
        gn = g.param()
        g.feedback_write(readers[n], g.node("ADD", g.node("MULTIPLY", f, gn), x))
        filt.append(f), gain.append(gn)
    g.output(sum_l, sum_r)
    return g, {"delay": d, "filter": filt, "gain": gain, "zero": [zero], "k": [k]}


def graph_chain256(n_nodes: int = 256) -> GraphSpec:
    """Config 5: NoiseGen feeding a chain cycling 8 node kinds.

    {multiply by const, add const, OnePole, Lopass, sinApprox, clamp(-1,1), abs,
    lerp with previous node}.  PARAM nodes carry the constants.
    """
    g = GraphSpec()
    k_mul, k_add, k_lo, k_hi, k_mix = g.param(), g.param(), g.param(), g.param(), g.param()
    x = g.node("NOISE")
    prev = x
    count = 0
    kinds = ["MULTIPLY", "ADD", "ONEPOLE", "LOPASS", "SIN_APPROX", "CLAMP", "ABS", "LERP"]
    while count < n_nodes:
        kind = kinds[count % 8]
        if kind == "MULTIPLY":
            y = g.node("MULTIPLY", x, k_mul)
        elif kind == "ADD":
            y = g.node("ADD", x, k_add)
        elif kind == "CLAMP":
            y = g.node("CLAMP", x, k_lo, k_hi)
        elif kind == "LERP":
            y = g.node("LERP", x, prev, k_mix)
        else:
            y = g.node(kind, x)
        prev, x = x, y
        count += 1
    return g.output(x)


# Aaltoverb (examples/audio-and-midi/reverb.cpp:21-123): allpass gains, setMaxDelayInSamples
# arguments and delay-time scales of the ten Allpass<PitchbendableDelay> sections and the two
# PitchbendableDelay feedback lines.
AALTOVERB_AP_GAINS = (0.75, 0.70, 0.625, 0.625, 0.7, 0.7, 0.6, 0.6, 0.5, 0.5)
AALTOVERB_AP_MAX = (500.0, 500.0, 1000.0, 1000.0, 2600.0, 2600.0, 8000.0, 8000.0, 10000.0, 10000.0)
AALTOVERB_AP_SCALE = (0.00476, 0.00358, 0.00973, 0.00830, 0.029, 0.021, 0.078, 0.090, 0.111, 0.096)
AALTOVERB_DELAY_MAX = 3500.0
AALTOVERB_DELAY_SCALE = (0.0313, 0.0371)


def graph_aaltoverb():
    """The reverb of examples/audio-and-midi/reverb.cpp:68-123 as a voice graph (one reverb per voice).

    Returns (spec, names) where names maps a label to its node index:
      PARAM nodes  'size2' (sizeU*2, the LinearGlide target), 'feedback', 'sr', 'vmin' (64), 'zero',
                   'apscale0..9', 'dscaleL', 'dscaleR'
      delay nodes  'ap1..ap10' (ALLPASS_PB), 'delayL', 'delayR' (PITCHBEND_DELAY),
                   'glideDelay', 'glideFeedback' (GLIDE)
    Inputs: planes 0 and 1 (ctx->inputs[0], [1]); outputs: vTapL, vTapR.
    """
    g = GraphSpec()
    n = {}
    in0, in1 = g.input(0), g.input(1)
    for key in ("size2", "feedback", "sr", "vmin", "zero"):
        n[key] = g.param()
    for i in range(10):
        n["apscale%d" % i] = g.param()
    n["dscaleL"], n["dscaleR"] = g.param(), g.param()
    n["glideDelay"] = g.node("GLIDE", n["size2"])          # reverb.cpp:86
    n["glideFeedback"] = g.node("GLIDE", n["feedback"])    # :87
    dparam = g.node("MULTIPLY", n["sr"], n["glideDelay"])  # :93
    mono = g.node("ADD", in0, in1)                         # :106

    def ap(i, x):  # r->mAp<i+1>(x, vt<i+1>), vt<i+1> = max(scale * delayParamInSamples, vMin) (:94-103);
        # the stateless vt row is computed next to its only reader so that few rows are live at once
        vt = g.node("MAX", g.node("MULTIPLY", n["apscale%d" % i], dparam), n["vmin"])
        n["ap%d" % (i + 1)] = g.node("ALLPASS_PB", x, vt)
        return n["ap%d" % (i + 1)]

    diffused = ap(3, ap(2, ap(1, ap(0, mono))))            # :107
    dtl = g.node("MAX", g.node("SUBTRACT", g.node("MULTIPLY", n["dscaleL"], dparam), n["vmin"]), n["zero"])  # :110
    dtr = g.node("MAX", g.node("SUBTRACT", g.node("MULTIPLY", n["dscaleR"], dparam), n["vmin"]), n["zero"])  # :111
    fbl, fbr = g.feedback_read(), g.feedback_read()        # mvFeedbackL, mvFeedbackR (:34)
    n["delayL"] = g.node("PITCHBEND_DELAY", fbl, dtl)
    n["delayR"] = g.node("PITCHBEND_DELAY", fbr, dtr)
    tap_l = ap(6, ap(4, g.node("ADD", diffused, n["delayL"])))   # :114
    tap_r = ap(7, ap(5, g.node("ADD", diffused, n["delayR"])))   # :115
    g.feedback_write(fbr, g.node("MULTIPLY", ap(8, tap_l), n["glideFeedback"]))   # :118
    g.feedback_write(fbl, g.node("MULTIPLY", ap(9, tap_r), n["glideFeedback"]))   # :119
    g.output(tap_l, tap_r)
    return g, n
