"""Shared helpers for the parity tests."""
import numpy as np


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_same_bits(got, want, what=""):
    """Bit-exact comparison; NaNs compare equal to NaNs (x86 and sm_100 canonical NaN
    payloads differ: 0xFFC00000 vs 0x7FFFFFFF -- both are 'a NaN')."""
    got = np.ascontiguousarray(got, np.float32)
    want = np.ascontiguousarray(want, np.float32)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    gb, wb = got.view(np.uint32), want.view(np.uint32)
    both_nan = np.isnan(got) & np.isnan(want)
    bad = (gb != wb) & ~both_nan
    if bad.any():
        idx = np.argwhere(bad)
        i = tuple(idx[0])
        raise AssertionError(
            f"{what}: {int(bad.sum())} of {bad.size} words differ; first at {i}: "
            f"got {got[i]!r} ({gb[i]:#010x}) want {want[i]!r} ({wb[i]:#010x})")


def assert_state_equal(got, want, what=""):
    got = np.ascontiguousarray(got, np.uint32)
    want = np.ascontiguousarray(want, np.uint32)
    gf, wf = got.view(np.float32), want.view(np.float32)
    both_nan = np.isnan(gf) & np.isnan(wf)
    bad = (got != want) & ~both_nan
    assert not bad.any(), f"{what}: state words differ at {np.argwhere(bad)[:4].tolist()}"


def run_gpu(api, w, n_blocks, inp, flags=0, want_mix=False, splits=None):
    """Run a Workload on the GPU through the host C-ABI entry point."""
    g = api.VoiceGraph(w.spec, w.n_voices, flags)
    try:
        g.set_coefs(w.coef)
        g.set_state(w.state)
        outs, mixes = [], []
        t0 = 0
        for n in (splits or (n_blocks,)):
            i = None if inp is None else np.ascontiguousarray(inp[t0:t0 + n])
            o, m = g.process_host(i, n, want_out=True, want_mix=want_mix)
            outs.append(o)
            mixes.append(m)
            t0 += n
        st = g.get_state()
        name = g.kernel_name
    finally:
        g.close()
    out = np.concatenate(outs, 0)
    mix = np.concatenate(mixes, 0) if want_mix else None
    return out, mix, st, name
