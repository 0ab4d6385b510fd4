"""ctypes bindings for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

  RefOracle   oracle/_ref/libmlref.so  : the unmodified reference (madronalib SSE
              path) compiled in place from /root/reference (oracle/Makefile `ref`).
  PortOracle  oracle/_port/libmlport.so: our plain-C restatement (oracle/port/mlport.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Nothing under madronalib_b200/ does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

from madronalib_b200.graph import BLOCK, GraphSpec

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libmlref.so")
E2S_LIB = os.path.join(_HERE, "_ref", "libmle2s.so")
PORT_LIB = os.path.join(_HERE, "_port", "libmlport.so")

_vp = ctypes.c_void_p


def build(target: str = "all") -> None:
    """Compile the checkers (port always; ref only where /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, target], check=True, stdout=subprocess.DEVNULL)


def ref_available() -> bool:
    return os.path.exists(REF_LIB)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_vp)


class _Oracle:
    prefix = ""
    path = ""

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(f"{self.path} not built; run `make -C oracle`")
        self.lib = ctypes.CDLL(self.path)
        p = self.prefix
        L = self.lib
        getattr(L, p + "graph_create").restype = _vp
        getattr(L, p + "graph_create").argtypes = [_vp, ctypes.c_int, _vp, ctypes.c_int,
                                                     ctypes.c_int, _vp]
        getattr(L, p + "graph_destroy").argtypes = [_vp]
        getattr(L, p + "graph_set_state").argtypes = [_vp, _vp]
        getattr(L, p + "graph_get_state").argtypes = [_vp, _vp]
        for name, nargs in (("lopass", 2), ("hipass", 2), ("bandpass", 2), ("loshelf", 3),
                            ("hishelf", 3), ("bell", 3), ("onepole", 1), ("peak", 1), ("rms", 1),
                            ("adsr", 5), ("glide", 1), ("sample_glide", 1)):
            fn = getattr(L, p + "coeffs_" + name)
            fn.argtypes = [ctypes.c_float] * nargs + [_vp]
            fn.restype = None
        getattr(L, p + "coeffs_dcblocker").argtypes = [ctypes.c_float]
        getattr(L, p + "coeffs_dcblocker").restype = ctypes.c_float
        getattr(L, p + "db_to_gain").argtypes = [ctypes.c_float]
        getattr(L, p + "db_to_gain").restype = ctypes.c_float
        getattr(L, p + "coeffs_allpass1").argtypes = [ctypes.c_float]
        getattr(L, p + "coeffs_allpass1").restype = ctypes.c_float

    # -- coefficient design --
    def coeffs_lopass_vec(self, omega: np.ndarray, k: np.ndarray) -> np.ndarray:
        """Lopass::makeCoeffsVec for one block: omega[64], k[64] -> [3][64]."""
        om, kk = np.ascontiguousarray(omega, np.float32), np.ascontiguousarray(k, np.float32)
        out = np.empty((3, BLOCK), np.float32)
        f = getattr(self.lib, self.prefix + "coeffs_lopass_vec")
        f.argtypes, f.restype = [_vp, _vp, _vp], None
        f(_ptr(om), _ptr(kk), _ptr(out))
        return out

    def coeffs(self, kind: str, *args: float) -> np.ndarray:
        n = {"lopass": 3, "hipass": 4, "bandpass": 3, "loshelf": 5, "hishelf": 6, "bell": 4,
             "onepole": 2, "peak": 2, "rms": 2, "adsr": 4, "glide": 2, "sample_glide": 2}[kind]
        out = np.zeros(n, np.float32)
        getattr(self.lib, self.prefix + "coeffs_" + kind)(*[ctypes.c_float(a) for a in args],
                                                          _ptr(out))
        return out

    def coeffs_dcblocker(self, omega: float) -> float:
        return float(getattr(self.lib, self.prefix + "coeffs_dcblocker")(omega))

    def coeffs_allpass1(self, d: float) -> float:
        return float(getattr(self.lib, self.prefix + "coeffs_allpass1")(d))

    def db_to_gain(self, db: float) -> float:
        return float(getattr(self.lib, self.prefix + "db_to_gain")(db))

    # -- graphs --
    def _process(self, h, inp, out, mix, T, nthreads, mix_mode, n_shards):
        raise NotImplementedError

    def run(self, spec: GraphSpec, n_voices: int, n_blocks: int, inp: Optional[np.ndarray],
            state: np.ndarray, coef: np.ndarray, want_out: bool = True, want_mix: bool = False,
            nthreads: int = 1, mix_mode: int = 0, n_shards: int = 1,
            splits: Optional[Tuple[int, ...]] = None):
        """Run T blocks from `state`; returns (out, mix, state_after).

        inp [T][n_in][V][64] f32; out [T][n_out][V][64]; mix [T][n_out][64].
        `splits`: process in several successive calls of these block counts
        (state carried inside the oracle object), to test launch-boundary continuity.
        """
        V, T = n_voices, n_blocks
        p = self.prefix
        coef = np.ascontiguousarray(coef, np.float32)
        state = np.ascontiguousarray(state, np.uint32)
        assert coef.shape == (spec.n_coef, V) and state.shape == (spec.n_state, V)
        if spec.n_in:
            inp = np.ascontiguousarray(inp, np.float32)
            assert inp.shape == (T, spec.n_in, V, BLOCK), inp.shape
        nodes, outs = spec.c_nodes(), spec.c_outs()
        h = getattr(self.lib, p + "graph_create")(nodes, spec.n_nodes, outs, spec.n_out, V,
                                                  _ptr(coef) if coef.size else None)
        if not h:
            raise RuntimeError("oracle rejected graph")
        try:
            if state.size:
                getattr(self.lib, p + "graph_set_state")(h, _ptr(state))
            out = np.zeros((T, spec.n_out, V, BLOCK), np.float32) if want_out else None
            mix = np.zeros((T, spec.n_out, BLOCK), np.float32) if want_mix else None
            t0 = 0
            for n in (splits or (T,)):
                i = inp[t0:t0 + n] if spec.n_in else None
                o = out[t0:t0 + n] if want_out else None
                m = mix[t0:t0 + n] if want_mix else None
                self._process(h, _ptr(i), _ptr(o), _ptr(m), n, nthreads, mix_mode, n_shards)
                t0 += n
            assert t0 == T
            st = np.zeros_like(state)
            if state.size:
                getattr(self.lib, p + "graph_get_state")(h, _ptr(st))
        finally:
            getattr(self.lib, p + "graph_destroy")(h)
        return out, mix, st


class RefOracle(_Oracle):
    prefix = "mlref_"
    path = REF_LIB

    def __init__(self):
        super().__init__()
        self.lib.mlref_graph_process.argtypes = [_vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int]
        self.lib.mlref_chain_sine_lopass_gain.restype = ctypes.c_double
        self.lib.mlref_chain_sine_lopass_gain.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp,
                                                          _vp, _vp, _vp, ctypes.c_int, ctypes.c_int]

        self.lib.mlref_aaltoverb.restype = ctypes.c_double
        self.lib.mlref_aaltoverb.argtypes = [ctypes.c_int, _vp, _vp, ctypes.c_float, ctypes.c_float,
                                             ctypes.c_float, ctypes.c_int]

    def shelf_vcoeffs(self, kind: str, p0, p1) -> np.ndarray:
        """LoShelf / HiShelf::vcoeffs({omega,k,A}, {omega,k,A}) -> [5 or 6][64]."""
        a, b = np.ascontiguousarray(p0, np.float32), np.ascontiguousarray(p1, np.float32)
        out = np.empty((5 if kind == "loshelf" else 6, BLOCK), np.float32)
        f = getattr(self.lib, "mlref_%s_vcoeffs" % kind)
        f.argtypes, f.restype = [_vp, _vp, _vp], None
        f(_ptr(a), _ptr(b), _ptr(out))
        return out

    def lopass_mod(self, x: np.ndarray, omega: np.ndarray, k: np.ndarray) -> np.ndarray:
        """Lopass::operator()(vx, omega, k) for ONE voice from cleared state; all [T][64]."""
        x, omega, k = (np.ascontiguousarray(a, np.float32) for a in (x, omega, k))
        out = np.empty_like(x)
        f = self.lib.mlref_lopass_mod
        f.argtypes, f.restype = [ctypes.c_int, _vp, _vp, _vp, _vp], None
        f(x.shape[0], _ptr(x), _ptr(omega), _ptr(k), _ptr(out))
        return out

    def kitchen(self, inp: np.ndarray) -> np.ndarray:
        """tests/cpp/kitchen_body.h compiled against the reference (one instance); inp [T][2][64] -> [T][2][64]."""
        inp = np.ascontiguousarray(inp, np.float32)
        out = np.empty_like(inp)
        self.lib.mlref_kitchen.argtypes = [ctypes.c_int, _vp, _vp]
        self.lib.mlref_kitchen.restype = None
        self.lib.mlref_kitchen(inp.shape[0], _ptr(inp), _ptr(out))
        return out

    def fdn(self, size: int, inp: np.ndarray, times, cutoffs, gains) -> np.ndarray:
        """The reference's FDN<size> object itself (size 4, 6, 8 or 16), one voice; inp [T][64] -> [T][2][64]."""
        inp = np.ascontiguousarray(inp, np.float32)
        t, c, g = (np.ascontiguousarray(a, np.float32) for a in (times, cutoffs, gains))
        out = np.empty((inp.shape[0], 2, 64), np.float32)
        self.lib.mlref_fdn_run.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]
        self.lib.mlref_fdn_run.restype = ctypes.c_int
        assert self.lib.mlref_fdn_run(size, inp.shape[0], _ptr(inp), _ptr(out), _ptr(t), _ptr(c), _ptr(g)) == 0
        return out

    def oversample_body(self, inp: np.ndarray) -> np.ndarray:
        """tests/cpp/oversample_body.h compiled against the reference (one instance); inp [T][1][64] -> [T][2][64]."""
        inp = np.ascontiguousarray(inp, np.float32)
        out = np.empty((inp.shape[0], 2, 64), np.float32)
        self.lib.mlref_oversample_body.argtypes = [ctypes.c_int, _vp, _vp]
        self.lib.mlref_oversample_body.restype = None
        self.lib.mlref_oversample_body(inp.shape[0], _ptr(inp), _ptr(out))
        return out

    def rest_body(self, inp: np.ndarray) -> np.ndarray:
        """tests/cpp/rest_body.h compiled against the reference (one instance); inp [T][2][64] -> [T][2][64]."""
        inp = np.ascontiguousarray(inp, np.float32)
        out = np.empty((inp.shape[0], 2, 64), np.float32)
        self.lib.mlref_rest_body.argtypes = [ctypes.c_int, _vp, _vp]
        self.lib.mlref_rest_body.restype = None
        self.lib.mlref_rest_body(inp.shape[0], _ptr(inp), _ptr(out))
        return out

    def rows_body(self, inp: np.ndarray) -> np.ndarray:
        """tests/cpp/rows_body.h compiled against the reference (one instance); inp [T][1][64] -> [T][2][64]."""
        inp = np.ascontiguousarray(inp, np.float32)
        out = np.empty((inp.shape[0], 2, 64), np.float32)
        self.lib.mlref_rows_body.argtypes = [ctypes.c_int, _vp, _vp]
        self.lib.mlref_rows_body.restype = None
        self.lib.mlref_rows_body(inp.shape[0], _ptr(inp), _ptr(out))
        return out

    def fdn_body(self, inp: np.ndarray) -> np.ndarray:
        """tests/cpp/fdn_body.h compiled against the reference (one instance); inp [T][2][64] -> [T][2][64]."""
        inp = np.ascontiguousarray(inp, np.float32)
        out = np.empty((inp.shape[0], 2, 64), np.float32)
        self.lib.mlref_fdn_body.argtypes = [ctypes.c_int, _vp, _vp]
        self.lib.mlref_fdn_body.restype = None
        self.lib.mlref_fdn_body(inp.shape[0], _ptr(inp), _ptr(out))
        return out

    def upsample_body(self, inp: np.ndarray) -> np.ndarray:
        """tests/cpp/upsample_body.h compiled against the reference (one instance); inp [T][2][64] -> [T][2][64]."""
        inp = np.ascontiguousarray(inp, np.float32)
        out = np.empty((inp.shape[0], 2, 64), np.float32)
        self.lib.mlref_upsample_body.argtypes = [ctypes.c_int, _vp, _vp]
        self.lib.mlref_upsample_body.restype = None
        self.lib.mlref_upsample_body(inp.shape[0], _ptr(inp), _ptr(out))
        return out

    def upsample2x_clip(self, inp: np.ndarray, drive: float) -> np.ndarray:
        """Upsample2xFunction<1> with fn = clamp(v * drive, -1, 1) for ONE voice; inp [T][64]."""
        inp = np.ascontiguousarray(inp, np.float32)
        out = np.empty_like(inp)
        self.lib.mlref_upsample2x_clip.argtypes = [ctypes.c_int, _vp, _vp, ctypes.c_float]
        self.lib.mlref_upsample2x_clip.restype = None
        self.lib.mlref_upsample2x_clip(inp.shape[0], _ptr(inp), _ptr(out), drive)
        return out

    def upsample2x_osc(self, inp: np.ndarray, phase0: int, g3) -> np.ndarray:
        """Upsample2xFunction<1> with the STATEFUL fn = lp(osc(v * 0.5)) (the reference's SineGen and Lopass called
        twice per vector) for ONE voice; inp [T][64]."""
        inp = np.ascontiguousarray(inp, np.float32)
        g3 = np.ascontiguousarray(g3, np.float32)
        out = np.empty_like(inp)
        self.lib.mlref_upsample2x_osc.argtypes = [ctypes.c_int, _vp, _vp, ctypes.c_uint32, _vp]
        self.lib.mlref_upsample2x_osc.restype = None
        self.lib.mlref_upsample2x_osc(inp.shape[0], _ptr(inp), _ptr(out), int(phase0), _ptr(g3))
        return out

    def downsample2x_clip(self, inp: np.ndarray, drive: float) -> np.ndarray:
        """Downsample2xFunction<1> with fn = clamp(v * drive, -1, 1) for ONE voice; inp [T][64]."""
        inp = np.ascontiguousarray(inp, np.float32)
        out = np.empty_like(inp)
        self.lib.mlref_downsample2x_clip.argtypes = [ctypes.c_int, _vp, _vp, ctypes.c_float]
        self.lib.mlref_downsample2x_clip.restype = None
        self.lib.mlref_downsample2x_clip(inp.shape[0], _ptr(inp), _ptr(out), drive)
        return out

    def aaltoverb(self, inp: np.ndarray, size_u2: float, feedback: float, glide_samples: float,
                  repeats: int = 1):
        """The reverb example's own per-vector body for ONE reverb; inp [T][2][64] -> (out, seconds)."""
        inp = np.ascontiguousarray(inp, np.float32)
        out = np.empty_like(inp)
        sec = self.lib.mlref_aaltoverb(inp.shape[0], _ptr(inp), _ptr(out), size_u2, feedback,
                                       glide_samples, repeats)
        return out, float(sec)

    def _process(self, h, inp, out, mix, T, nthreads, mix_mode, n_shards):
        if mix is not None and mix_mode != 0:
            raise ValueError("the reference only sums voices left to right (mix_mode 0)")
        self.lib.mlref_graph_process(h, inp, out, mix, T, nthreads)

    def sizeof(self, which: int) -> int:
        return int(self.lib.mlref_sizeof(which))

    def chain_sine_lopass_gain(self, inp: np.ndarray, coef3: np.ndarray, gain: np.ndarray,
                               phase: np.ndarray, ic: np.ndarray, nthreads: int, repeats: int = 1):
        """The reference's own chain loop (struct Voice{SineGen; Lopass}); returns (out, seconds)."""
        T, V, _ = inp.shape
        out = np.empty_like(inp)
        sec = self.lib.mlref_chain_sine_lopass_gain(V, T, _ptr(inp), _ptr(out), _ptr(coef3),
                                                    _ptr(gain), _ptr(phase), _ptr(ic), nthreads, repeats)
        return out, float(sec)


class PortOracle(_Oracle):
    prefix = "mlport_"
    path = PORT_LIB

    def __init__(self):
        super().__init__()
        self.lib.mlport_graph_process.argtypes = [_vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_int]

    def _process(self, h, inp, out, mix, T, nthreads, mix_mode, n_shards):
        self.lib.mlport_graph_process(h, inp, out, mix, T, nthreads, mix_mode, n_shards)


class _VoiceBank:
    """EventsToSignals::Voice x V.  prefix/lib as for the graph oracles."""

    def __init__(self, lib, prefix):
        self.lib, self.p = lib, prefix
        getattr(lib, prefix + "bank_create").restype = _vp
        getattr(lib, prefix + "bank_create").argtypes = [ctypes.c_int, ctypes.c_float, _vp, _vp, _vp, _vp,
                                                         ctypes.c_uint]
        getattr(lib, prefix + "bank_destroy").argtypes = [_vp]
        getattr(lib, prefix + "bank_process").restype = ctypes.c_double
        getattr(lib, prefix + "bank_process").argtypes = [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int]

    def run(self, sr, voice_index, glide_seconds, drift_amount, pitch_bend, events, splits=None, nthreads=1,
            flags=0):
        """events [T][V] records -> (out [T][8][V][64], seconds)."""
        T, V = events.shape
        vi = np.ascontiguousarray(voice_index, np.int32)
        gs, da, pb = (np.ascontiguousarray(a, np.float32) for a in (glide_seconds, drift_amount, pitch_bend))
        h = getattr(self.lib, self.p + "bank_create")(V, sr, _ptr(vi), _ptr(gs), _ptr(da), _ptr(pb), flags)
        out = np.zeros((T, 8, V, BLOCK), np.float32)
        sec, t0 = 0.0, 0
        try:
            for n in (splits or (T,)):
                e = np.ascontiguousarray(events[t0:t0 + n])
                sec += getattr(self.lib, self.p + "bank_process")(h, n, _ptr(e), _ptr(out[t0:t0 + n]), nthreads)
                t0 += n
        finally:
            getattr(self.lib, self.p + "bank_destroy")(h)
        return out, float(sec)


def ref_voice_bank() -> _VoiceBank:
    """The reference's own EventsToSignals::Voice, compiled in place (oracle/_ref/libmle2s.so)."""
    if not os.path.exists(E2S_LIB):
        raise FileNotFoundError(E2S_LIB + " not built; run `make -C oracle ref`")
    return _VoiceBank(ctypes.CDLL(E2S_LIB), "mle2s_")


def port_voice_bank() -> _VoiceBank:
    if not os.path.exists(PORT_LIB):
        build("port")
    return _VoiceBank(ctypes.CDLL(PORT_LIB), "mlport_")


class Resampler:
    """Upsampler / Downsampler bank.  which = "ref" (the reference's own classes) or "port"."""

    def __init__(self, which: str, direction: int, octaves: int, n_voices: int):
        self.p = "mlref_" if which == "ref" else "mlport_"
        self.lib = ctypes.CDLL(REF_LIB if which == "ref" else PORT_LIB)
        f = getattr(self.lib, self.p + "resampler_create")
        f.restype, f.argtypes = _vp, [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        getattr(self.lib, self.p + "resampler_destroy").argtypes = [_vp]
        g = getattr(self.lib, self.p + "resampler_process")
        g.restype, g.argtypes = ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int]
        self.dir, self.oct, self.V = direction, octaves, n_voices
        self.h = f(direction, octaves, n_voices)

    def process(self, x: np.ndarray) -> np.ndarray:
        """x [T][V][64] -> [T_out][V][64]"""
        x = np.ascontiguousarray(x, np.float32)
        T = x.shape[0]
        cap = T << self.oct if self.dir == 0 else T
        out = np.zeros((max(cap, 1), self.V, BLOCK), np.float32)
        n = getattr(self.lib, self.p + "resampler_process")(self.h, _ptr(x), _ptr(out), T)
        return out[:n]

    def close(self):
        if self.h:
            getattr(self.lib, self.p + "resampler_destroy")(self.h)
            self.h = None


class RefEventsToSignals:
    """The complete reference EventsToSignals for one instrument (oracle/_ref/libmle2s.so)."""

    def __init__(self, sr: float, polyphony: int, glide_seconds: float, drift_amount: float, unison: bool = False,
                 mpe: bool = False):
        if not os.path.exists(E2S_LIB):
            raise FileNotFoundError(E2S_LIB + " not built; run `make -C oracle ref`")
        self.lib = L = ctypes.CDLL(E2S_LIB)
        L.mle2s_full_create.restype = _vp
        L.mle2s_full_create.argtypes = [ctypes.c_float, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                        ctypes.c_int]
        L.mle2s_full_destroy.argtypes = [_vp]
        L.mle2s_full_add_event.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_float, ctypes.c_float]
        L.mle2s_full_process.argtypes = [_vp, ctypes.c_int, _vp]
        self.P = polyphony
        self.h = L.mle2s_full_create(sr, polyphony, glide_seconds, drift_amount, int(unison), int(mpe))

    def add_event(self, type, channel, source_idx, time, value1=0.0, value2=0.0):
        self.lib.mle2s_full_add_event(self.h, type, channel, source_idx, time, value1, value2)

    def process_vector(self, start: int) -> np.ndarray:
        """-> [polyphony][8][64]"""
        out = np.zeros((self.P, 8, BLOCK), np.float32)
        self.lib.mle2s_full_process(self.h, start, _ptr(out))
        return out

    def close(self):
        if self.h:
            self.lib.mle2s_full_destroy(self.h)
            self.h = None
