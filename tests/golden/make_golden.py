#!/usr/bin/env python
"""Generate the committed golden vectors from the REFERENCE ITSELF (oracle/_ref/libmlref.so,
i.e. madronalib's own headers compiled in place).  Run only in the authoring container:

    make -C oracle ref && python tests/golden/make_golden.py

The .npz files travel with the repo so that the GPU box (which has no /root/reference) can
check both the C port and the CUDA kernels against outputs of the real reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from madronalib_b200 import workloads as wl  # noqa: E402
from madronalib_b200.graph import OP_TABLE, GraphSpec  # noqa: E402
from oracle.bindings import RefOracle  # noqa: E402

SPECIALS = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-40, -1e-40, 1e38, -1e38, 3e9, -3e9,
                     2147483648.0, -2147483648.0, 0.5, 1.5, 2.5, -0.5, -1.5, 88.5, -88.5, 100, -100,
                     1.0, -1.0, np.pi, -np.pi, 8191.5, 1e-20, 0.70710678, 1e-3, -1e-3, 7.0],
                    np.float32)


def stateless_inputs(nin, rows=6, seed=7):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((1, nin, rows, 64)) * 4.0).astype(np.float32)
    for k in range(nin):
        x[0, k, 0, :SPECIALS.size] = np.roll(SPECIALS, 3 * k)
        x[0, k, 1] = np.linspace(-np.pi, np.pi, 64, dtype=np.float32)
    return x


def main():
    R = RefOracle()
    assert [R.sizeof(i) for i in range(6)] == [256, 20, 12, 4, 48, 2560]
    out = {}

    # every stateless op
    for name, (_, nin, nst, nco) in OP_TABLE.items():
        if not (30 <= OP_TABLE[name][0] < 100):  # MLB_OP_MAP_FIRST..END
            continue
        g = GraphSpec()
        g.output(g.node(name, *[g.input(k) for k in range(nin)]))
        x = stateless_inputs(nin)
        y, _, _ = R.run(g, x.shape[2], 1, x, g.new_state(x.shape[2]), g.new_coefs(x.shape[2]))
        out[f"op_{name}_in"] = x
        out[f"op_{name}_out"] = y
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **out)

    # the BASELINE configurations at oracle-friendly sizes
    cfgs = {
        "config1": (wl.config_1(), 4),
        "configA": (wl.config_a(40), 3),
        "config2_hipass": (wl.config_2("hipass", 33), 3),
        "config2_bandpass": (wl.config_2("bandpass", 33), 3),
        "config2_loshelf": (wl.config_2("loshelf", 33), 3),
        "config2_hishelf": (wl.config_2("hishelf", 33), 3),
        "config2_bell": (wl.config_2("bell", 33), 3),
        "config3": (wl.config_3(40), 3),
        "config4": (wl.config_4(16), 12),
        "config5": (wl.config_5(8, 256), 2),
    }
    out = {}
    for name, (w, T) in cfgs.items():
        inp = w.inputs(T)
        if name == "configA":  # cvtps2dq overflow / NaN canaries (SURVEY hard part 2)
            inp[0, 0, 0, :4] = [0.6, 0.5, -0.3, np.nan]
        y, mix, st = R.run(w.spec, w.n_voices, T, inp, w.state, w.coef, want_mix=True)
        if inp is not None:
            out[name + "_in"] = inp
        out[name + "_coef"] = w.coef
        out[name + "_state0"] = w.state
        out[name + "_out"] = y
        out[name + "_mix"] = mix
        out[name + "_state1"] = st
        print(name, y.shape, float(np.nanmax(np.abs(y))))
    np.savez_compressed(os.path.join(HERE, "configs.npz"), **out)

    # coefficient design
    out = {}
    om = np.linspace(0.001, 0.45, 64, dtype=np.float32)
    for kind, extra in (("lopass", (0.5,)), ("hipass", (0.5,)), ("bandpass", (0.3,)),
                        ("loshelf", (0.7, 1.4)), ("hishelf", (0.7, 1.4)), ("bell", (0.5, 1.41)),
                        ("onepole", ())):
        out["coef_" + kind] = np.stack([R.coeffs(kind, float(o), *extra) for o in om])
    out["coef_omega"] = om
    out["db_to_gain_6"] = np.float32(R.db_to_gain(6.0))
    out["dcblocker_0045"] = np.float32(R.coeffs_dcblocker(0.045))
    np.savez_compressed(os.path.join(HERE, "coeffs.npz"), **out)
    make_functors(R)
    make_voices()
    for f in ("ops.npz", "configs.npz", "coeffs.npz", "functors.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


FUNCTOR_V, FUNCTOR_T = 6, 10
AALTOVERB_V, AALTOVERB_T = 2, 48


def make_functors(R):
    """every added functor and the Aaltoverb example chain, from the reference."""
    out = {}
    for name in wl.FUNCTOR_CASES + wl.AGAIN_CASES:
        w = wl.functor_case(name, FUNCTOR_V)
        inp = w.inputs(FUNCTOR_T)
        y, _, st = R.run(w.spec, w.n_voices, FUNCTOR_T, inp, w.state, w.coef)
        out[name + "_in"], out[name + "_coef"], out[name + "_state0"] = inp, w.coef, w.state
        out[name + "_out"], out[name + "_state1"] = y, st
    w = wl.config_6(AALTOVERB_V)
    inp = w.inputs(AALTOVERB_T)
    y, _, st = R.run(w.spec, w.n_voices, AALTOVERB_T, inp, w.state, w.coef)
    # the example's own per-vector body (mlref_aaltoverb) gives the same bits as the graph
    for v in range(AALTOVERB_V):
        o, _ = R.aaltoverb(inp[:, :, v, :], float(w.coef[w.spec.coef_slot(2), v]),
                           float(w.coef[w.spec.coef_slot(3), v]), 4800.0)
        assert np.array_equal(o.view(np.uint32), y[:, :, v, :].view(np.uint32))
    out["aaltoverb_in"], out["aaltoverb_coef"], out["aaltoverb_state0"] = inp, w.coef, w.state
    out["aaltoverb_out"], out["aaltoverb_state1"] = y, st
    out["coef_adsr"] = R.coeffs("adsr", 0.01, 0.1, 0.5, 0.2, 48000.0)
    out["coef_peak"] = R.coeffs("peak", 0.001)
    out["coef_rms"] = R.coeffs("rms", 0.002)
    out["coef_glide"] = R.coeffs("glide", 4800.0)
    out["coef_sample_glide"] = R.coeffs("sample_glide", 77.7)
    import ctypes
    tab = np.zeros(17, np.float32)
    R.lib.mlref_impulse_table(tab.ctypes.data_as(ctypes.c_void_p))
    out["impulse_table"] = tab
    out["coef_allpass1"] = np.array([R.coeffs_allpass1(float(d)) for d in np.linspace(0.618, 1.618, 16, dtype=np.float32)],
                                    np.float32)
    np.savez_compressed(os.path.join(HERE, "functors.npz"), **out)


def make_voices():
    """EventsToSignals::Voice rows from the reference's own Voice (libmle2s.so)."""
    from oracle.bindings import ref_voice_bank
    V, T, sr = 6, 40, 48000.0
    ev = wl.voice_events(V, T, seed=11)
    out, _ = ref_voice_bank().run(sr, *wl.voice_bank_params(V), ev, flags=wl.VOICES_MIDI)
    np.savez_compressed(os.path.join(HERE, "voices.npz"), events=ev.view(np.uint8), shape=np.array([T, V]),
                        sr=np.float32(sr), out=out)
    print("voices.npz", os.path.getsize(os.path.join(HERE, "voices.npz")), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "voices":
        make_voices()
    elif len(sys.argv) > 1 and sys.argv[1] == "functors":
        make_functors(RefOracle())
        print("functors.npz", os.path.getsize(os.path.join(HERE, "functors.npz")), "bytes")
    else:
        main()
