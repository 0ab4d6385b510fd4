"""Synthetic workloads for the BASELINE.json configurations.

Deterministic formulas only (no RNG except NoiseGen's own LCG).  Coefficients come from
the library's host-side ``mlb_coeffs_*`` (glibc libm, the same calls as the reference's
``makeCoeffs``); tests check those against both oracles.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import api
from .graph import (BLOCK, OP_ID, SINE_ZERO_PHASE, GraphSpec, graph_chain256, graph_fm3_fdn8,
                    graph_phasor_lopass_onepole, graph_sine_lopass_gain, graph_sine_svf)

SR = 48000.0


@dataclass
class Workload:
    name: str
    spec: GraphSpec
    n_voices: int
    coef: np.ndarray   # [n_coef][V] f32
    state: np.ndarray  # [n_state][V] u32

    def inputs(self, n_blocks: int, t0: int = 0, v0: int = 0, v1: Optional[int] = None,
               out: Optional[np.ndarray] = None) -> Optional[np.ndarray]:
        """Contract R input planes [T][n_in][V'][64] for voices [v0, v1), blocks [t0, t0+T)."""
        if self.spec.n_in == 0:
            return None
        v1 = self.n_voices if v1 is None else v1
        return freq_rows(self.n_voices, n_blocks, t0, v0, v1, out)

    def shard(self, rank: int, world: int) -> "Workload":
        """Contiguous voice range [rank*V/world, (rank+1)*V/world)."""
        v0, v1 = self.n_voices * rank // world, self.n_voices * (rank + 1) // world
        w = Workload(self.name, self.spec, v1 - v0, np.ascontiguousarray(self.coef[:, v0:v1]),
                     np.ascontiguousarray(self.state[:, v0:v1]))
        w._v0, w._V = v0, self.n_voices  # type: ignore[attr-defined]
        return w


def base_freq(n_voices: int, v0: int = 0, v1: Optional[int] = None) -> np.ndarray:
    """f_v = (110 + 0.25 v) / 48000 cycles per sample (max 0.344 at V = 65536)."""
    v1 = n_voices if v1 is None else v1
    v = np.arange(v0, v1, dtype=np.float32)
    return ((np.float32(110.0) + np.float32(0.25) * v) / np.float32(SR)).astype(np.float32)


def freq_rows(n_voices: int, n_blocks: int, t0: int = 0, v0: int = 0, v1: Optional[int] = None,
              out: Optional[np.ndarray] = None) -> np.ndarray:
    """Time-varying per-voice frequency rows (contract R):
    freq_v[n] = f_v * (1 + 0.002 * tri(((n*7 + v*13) mod 512) / 512)), n = absolute sample index,
    tri = 0..1..0 triangle.  Shape [T][1][V'][64] f32."""
    v1 = n_voices if v1 is None else v1
    nv = v1 - v0
    if out is None:
        out = np.empty((n_blocks, 1, nv, BLOCK), np.float32)
    f = base_freq(n_voices, v0, v1)[:, None]
    vv = (np.arange(v0, v1, dtype=np.int64) * 13)[:, None]
    for t in range(n_blocks):
        n = (np.arange(BLOCK, dtype=np.int64) + (t0 + t) * BLOCK)[None, :] * 7
        ph = ((n + vv) % 512).astype(np.float32) / np.float32(512.0)
        tri = np.float32(1.0) - np.abs(np.float32(2.0) * ph - np.float32(1.0))
        out[t, 0] = f * (np.float32(1.0) + np.float32(0.002) * tri)
    return out


def _set(coef: np.ndarray, spec: GraphSpec, node: int, values) -> None:
    values = np.asarray(values, np.float32)
    base = spec.coef_slot(node)
    if values.ndim == 1:
        values = values[:, None]
    coef[base:base + values.shape[0]] = values


def _svf_coefs(kind: str, n_voices: int) -> np.ndarray:
    """omega_v = 0.02 + 0.2 v/V, k = 0.5 (Bell/shelves: A = dBToGain(6)); [n_coef][V]."""
    n = api._NCOEF[kind]
    out = np.zeros((n, n_voices), np.float32)
    A = api.db_to_gain(6.0)
    for v in range(n_voices):
        omega = np.float32(0.02) + np.float32(0.2) * np.float32(v) / np.float32(n_voices)
        if kind in ("lopass", "hipass", "bandpass"):
            out[:, v] = api.coeffs(kind, float(omega), 0.5)
        else:
            out[:, v] = api.coeffs(kind, float(omega), 0.5, A)
    return out


def config_a(n_voices: int = 65536, gain: float = 0.1) -> Workload:
    """Headline: SineGen.clear() -> Lopass(fixed coeffs) -> gain (also config 1 at V=1)."""
    spec = graph_sine_lopass_gain()
    coef = spec.new_coefs(n_voices)
    _set(coef, spec, 2, _svf_coefs("lopass", n_voices))
    coef[spec.coef_slot(3)] = np.float32(gain)
    state = spec.new_state(n_voices)
    state[spec.state_slot(1)] = SINE_ZERO_PHASE
    return Workload("sine_lopass_gain", spec, n_voices, coef, state)


def config_1() -> Workload:
    """Plumbing: 1 voice, freq 440/48000 constant, Lopass::makeCoeffs(0.1, 1.0), gain 0.5."""
    w = config_a(1, gain=0.5)
    _set(w.coef, w.spec, 2, api.coeffs("lopass", 0.1, 1.0))
    w.name = "config1"
    w.inputs = lambda n_blocks, t0=0, v0=0, v1=None, out=None: np.full(  # type: ignore[assignment]
        (n_blocks, 1, 1, BLOCK), np.float32(440.0) / np.float32(48000.0), np.float32)
    return w


def config_2(kind: str = "lopass", n_voices: int = 4096) -> Workload:
    """SineGen -> SVF ("Biquad" stand-in)."""
    spec = graph_sine_svf(kind)
    coef = spec.new_coefs(n_voices)
    _set(coef, spec, 2, _svf_coefs(kind, n_voices))
    state = spec.new_state(n_voices)
    state[spec.state_slot(1)] = SINE_ZERO_PHASE
    return Workload("sine_" + kind, spec, n_voices, coef, state)


def config_3(n_voices: int = 65536) -> Workload:
    """PhasorGen.clear(0) -> Lopass -> OnePole(0.001 + 0.01 v/V)."""
    spec = graph_phasor_lopass_onepole()
    coef = spec.new_coefs(n_voices)
    _set(coef, spec, 2, _svf_coefs("lopass", n_voices))
    op = np.zeros((2, n_voices), np.float32)
    for v in range(n_voices):
        om = np.float32(0.001) + np.float32(0.01) * np.float32(v) / np.float32(n_voices)
        op[:, v] = api.coeffs("onepole", float(om))
    _set(coef, spec, 3, op)
    return Workload("phasor_lopass_onepole", spec, n_voices, coef, spec.new_state(n_voices))


FDN_TIMES = np.array([67, 73, 91, 103, 127, 151, 173, 199], np.float32)
FDN_CUTOFFS = np.array([0.1, 0.2, 0.3, 0.4, 0.1, 0.2, 0.3, 0.4], np.float32)


def config_4(n_voices: int = 16384) -> Workload:
    """3-op FM (r1=2, r2=3.5, i1=0.3, i2=0.1) -> FDN<8>, delays {67..199}+64*(v mod 8),
    cutoffs {0.1,0.2,0.3,0.4}x2, feedback 0.5, stereo out."""
    spec = graph_fm3_fdn8()
    coef = spec.new_coefs(n_voices)
    for node, val in zip(range(1, 6), (2.0, 3.5, 0.3, 0.1, 1.0)):
        coef[spec.coef_slot(node)] = np.float32(val)
    fdn = spec.ops.index(OP_ID["FDN8"])
    gains = np.full(8, 0.5, np.float32)
    per = [api.coeffs_fdn8(FDN_TIMES + np.float32(64 * m), FDN_CUTOFFS, gains) for m in range(8)]
    block = np.stack([per[v % 8] for v in range(n_voices)], axis=1)  # [32][V]
    _set(coef, spec, fdn, block)
    state = spec.new_state(n_voices)
    for i, op in enumerate(spec.ops):
        if op == OP_ID["SINE"]:
            state[spec.state_slot(i)] = SINE_ZERO_PHASE
    w = Workload("fm3_fdn8", spec, n_voices, coef, state)
    w.inputs = lambda n_blocks, t0=0, v0=0, v1=None, out=None: np.ascontiguousarray(  # type: ignore
        np.broadcast_to(base_freq(n_voices, v0, v1)[None, None, :, None],
                        (n_blocks, 1, (n_voices if v1 is None else v1) - v0, BLOCK)))
    return w


def config_5(n_instances: int = 1024, n_nodes: int = 256) -> Workload:
    """256-node chain cycling 8 node kinds, fed by NoiseGen seeded with the instance index."""
    spec = graph_chain256(n_nodes)
    coef = spec.new_coefs(n_instances)
    for node, val in zip(range(5), (0.999, 1e-3, -1.0, 1.0, 0.5)):
        coef[spec.coef_slot(node)] = np.float32(val)
    c_op, c_lp = api.coeffs("onepole", 0.01), api.coeffs("lopass", 0.1, 1.0)
    for i, op in enumerate(spec.ops):
        if op == OP_ID["ONEPOLE"]:
            _set(coef, spec, i, np.repeat(c_op[:, None], n_instances, 1))
        elif op == OP_ID["LOPASS"]:
            _set(coef, spec, i, np.repeat(c_lp[:, None], n_instances, 1))
    state = spec.new_state(n_instances)
    state[spec.state_slot(spec.ops.index(OP_ID["NOISE"]))] = np.arange(n_instances, dtype=np.uint32)
    return Workload("chain%d" % n_nodes, spec, n_instances, coef, state)
