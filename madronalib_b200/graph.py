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
MAX_INS = 3

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


OP_TABLE = _parse_op_table()
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
        for op in self.ops:
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
            for k in range(MAX_INS):
                arr[i].inp[k] = self.ins[i][k]
            arr[i].iarg = self.iargs[i]
        return arr

    def c_outs(self):
        return (ctypes.c_int32 * max(1, self.n_out))(*self.outs)

    def new_state(self, n_voices: int) -> np.ndarray:
        return np.zeros((max(1, self.n_state), n_voices), dtype=np.uint32)[: self.n_state]

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
