"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracles."""
import numpy as np
import pytest

from madronalib_b200 import workloads as wl
from madronalib_b200.graph import GraphSpec, OP_ID
from tests.common import assert_same_bits, assert_state_equal, run_gpu

pytestmark = pytest.mark.gpu


def test_config1_pinned_values(gpu):
    """SURVEY 8c spot values of the reference: 0.5*Lopass{0.1,1.0}(SineGen.clear()(440/48000))."""
    w = wl.config_1()
    out, _, _, name = run_gpu(gpu, w, 1, w.inputs(1))
    assert name.startswith("fused:")
    y = out[0, 0, 0]
    assert float(y[0]).hex() == "-0x1.09e6400000000p-9"
    assert float(y[1]).hex() == "-0x1.5cd8fa0000000p-7"
    assert float(y[63]).hex() == "0x1.b253900000000p-3"


def test_sinegen_one_cycle_ends_at_zero(gpu):
    """Reference Tests/dspGensTest.cpp:22-31: |SineGen(1/64)[63]| < dBToAmp(-120)."""
    g = GraphSpec()
    g.output(g.node("SINE", g.input(0)))
    w = wl.Workload("sine", g, 1, g.new_coefs(1), g.new_state(1))
    w.state[0] = wl.SINE_ZERO_PHASE
    inp = np.full((1, 1, 1, 64), 1.0 / 64, np.float32)
    out, _, _, _ = run_gpu(gpu, w, 1, inp)
    assert abs(out[0, 0, 0, 63]) < 10.0 ** (-120 / 20.0)
    assert float(out[0, 0, 0, 63]).hex() == "-0x1.0f876c0000000p-22"


@pytest.mark.parametrize("n_voices,n_blocks", [(1, 3), (31, 2), (32, 4), (200, 5), (1000, 3)])
def test_config_a_bit_exact(gpu, port, n_voices, n_blocks):
    w = wl.config_a(n_voices)
    inp = w.inputs(n_blocks)
    inp[0, 0, 0, :4] = [0.6, 0.5, -0.3, np.nan] if n_voices >= 1 else 0  # cvtps2dq overflow canaries
    po, _, ps = port.run(w.spec, n_voices, n_blocks, inp, w.state, w.coef)
    go, _, gs, name = run_gpu(gpu, w, n_blocks, inp)
    assert name.startswith("fused:"), name
    assert_same_bits(go, po, "config A out")
    assert_state_equal(gs, ps, "config A state")


def test_config_a_against_reference_itself(gpu, ref):
    w = wl.config_a(96)
    inp = w.inputs(6)
    ro, _, rs = ref.run(w.spec, 96, 6, inp, w.state, w.coef)
    go, _, gs, _ = run_gpu(gpu, w, 6, inp)
    assert_same_bits(go, ro, "config A vs reference")
    assert_state_equal(gs, rs)


@pytest.mark.parametrize("n_voices,n_blocks,chunks,warps", [(100, 24, None, None), (333, 7, 4, None),
                                                            (96, 9, 9, 3), (2100, 16, 2, 2)])
def test_dynamic_work_units(gpu, port, monkeypatch, n_voices, n_blocks, chunks, warps):
    """The launch is cut into (group, chunk) work units pulled from an atomic queue; a voice's
    state hops between warps at chunk boundaries.  Any chunking must give identical bits."""
    if chunks is not None:
        monkeypatch.setenv("MLB_CHAIN_CHUNKS", str(chunks))
    if warps is not None:
        monkeypatch.setenv("MLB_CHAIN_WARPS", str(warps))
    w = wl.config_a(n_voices)
    inp = w.inputs(n_blocks)
    po, pm, ps = port.run(w.spec, n_voices, n_blocks, inp, w.state, w.coef, want_mix=True, mix_mode=1,
                          nthreads=4)
    go, gm, gs, name = run_gpu(gpu, w, n_blocks, inp, want_mix=True)
    assert name.startswith("fused:"), name
    assert_same_bits(go, po, "out")
    assert_same_bits(gm, pm, "mix")
    assert_state_equal(gs, ps)
    # and again in two launches (progress counters carry over)
    go2, _, gs2, _ = run_gpu(gpu, w, n_blocks, inp, splits=(n_blocks // 2, n_blocks - n_blocks // 2))
    assert_same_bits(go2, po, "out (two launches)")
    assert_state_equal(gs2, ps)


@pytest.mark.parametrize("slices", [1, 5, 8])
def test_host_entry_point_voice_slices(gpu, port, monkeypatch, slices):
    """mlb_graph_process_host pipelines big banks over voice slices (H2D / kernel / D2H on three
    streams, cudaMemcpy2DAsync windows, per-slice progress offsets); the result must not depend on
    the slicing.  The slicing threshold is lowered so that this bank really takes the pipeline
    (asserted through mlb_graph_last_host_slices); the full-size case is in test_gpu_fullsize.py."""
    monkeypatch.setenv("MLB_HOST_SLICES", str(slices))
    monkeypatch.setenv("MLB_HOST_SLICE_MIN_MB", "1")
    V, T = 4200 + 7, 17
    w = wl.config_a(V)
    inp = w.inputs(T)
    po, pm, ps = port.run(w.spec, V, T, inp, w.state, w.coef, want_mix=True, mix_mode=1, nthreads=8)
    g = gpu.VoiceGraph(w.spec, V)
    try:
        g.set_coefs(w.coef)
        g.set_state(w.state)
        outs, mixes = [], []
        for t0, n in ((0, 9), (9, 8)):
            o, m = g.process_host(np.ascontiguousarray(inp[t0:t0 + n]), n, want_out=True, want_mix=True)
            # ceil(V / slices) rounded up to whole 32-voice groups per slice
            per = -(-(-(-V // slices)) // 32) * 32
            assert g.last_host_slices == (-(-V // per) if slices > 1 else 1), g.last_host_slices
            outs.append(o)
            mixes.append(m)
        gs = g.get_state()
    finally:
        g.close()
    assert_same_bits(np.concatenate(outs), po, "out")
    assert_same_bits(np.concatenate(mixes), pm, "mix")
    assert_state_equal(gs, ps)


def test_launch_boundary_continuity(gpu, port):
    """State carried across launches: 2+1+4 blocks == 7 blocks."""
    w = wl.config_a(100)
    inp = w.inputs(7)
    po, _, ps = port.run(w.spec, 100, 7, inp, w.state, w.coef)
    go, _, gs, _ = run_gpu(gpu, w, 7, inp, splits=(2, 1, 4))
    assert_same_bits(go, po)
    assert_state_equal(gs, ps)


@pytest.mark.parametrize("kind", ["lopass", "hipass", "bandpass", "loshelf", "hishelf", "bell"])
def test_config2_svf_family(gpu, port, kind):
    w = wl.config_2(kind, 160)
    inp = w.inputs(4)
    po, _, ps = port.run(w.spec, 160, 4, inp, w.state, w.coef)
    go, _, gs, name = run_gpu(gpu, w, 4, inp)
    assert name.startswith("fused:"), name
    assert_same_bits(go, po, kind)
    assert_state_equal(gs, ps, kind)


def test_config3_phasor_lopass_onepole(gpu, port):
    w = wl.config_3(300)
    inp = w.inputs(5)
    po, _, ps = port.run(w.spec, 300, 5, inp, w.state, w.coef)
    go, _, gs, name = run_gpu(gpu, w, 5, inp)
    assert name.startswith("fused:"), name
    assert_same_bits(go, po)
    assert_state_equal(gs, ps)


def test_generic_kernel_equals_fused(gpu, port):
    w = wl.config_a(70)
    inp = w.inputs(3)
    fo, _, fs, fname = run_gpu(gpu, w, 3, inp)
    go, _, gs, gname = run_gpu(gpu, w, 3, inp, flags=gpu.FLAG_FORCE_GENERIC)
    assert fname.startswith("fused:") and gname.startswith("generic")
    assert_same_bits(go, fo)
    assert_state_equal(gs, fs)


def test_mix_bus(gpu, port):
    """Mix bus = sum over voices.  Bit-exact against the port in the device summation order
    (32-voice groups left to right, groups left to right); within (V-1)*eps*sum|x_v| of the
    reference's strict left-to-right order (SURVEY hard part 8)."""
    V, T = 200, 3
    w = wl.config_a(V)
    inp = w.inputs(T)
    po, pm1, _ = port.run(w.spec, V, T, inp, w.state, w.coef, want_mix=True, mix_mode=1)
    _, pm0, _ = port.run(w.spec, V, T, inp, w.state, w.coef, want_mix=True, mix_mode=0)
    for flags in (0, gpu.FLAG_FORCE_GENERIC):
        go, gm, _, _ = run_gpu(gpu, w, T, inp, flags=flags, want_mix=True)
        assert_same_bits(go, po)
        assert_same_bits(gm, pm1, "mix (device order)")
        tol = V * np.finfo(np.float32).eps * np.abs(po).sum(axis=2).max()  # (n-1) eps sum|x_i|
        assert np.abs(gm - pm0).max() <= tol


def test_mix_bus_many_chunks(gpu, port):
    """More than one 2048-voice chunk and a ragged last group: exercises every level of the
    deterministic mix-bus tree (32-voice groups -> 64-group chunks -> total)."""
    V, T = 2 * 2048 + 100, 2
    w = wl.config_a(V)
    inp = w.inputs(T)
    po, pm1, _ = port.run(w.spec, V, T, inp, w.state, w.coef, want_mix=True, mix_mode=1, nthreads=8)
    go, gm, _, _ = run_gpu(gpu, w, T, inp, want_mix=True)
    assert_same_bits(go, po)
    assert_same_bits(gm, pm1, "mix (device order, 3 chunks)")


def test_noise_chain_no_input(gpu, port):
    g = GraphSpec()
    n = g.node("NOISE")
    lp = g.node("LOPASS", n)
    k = g.param()
    g.output(g.node("MULTIPLY", lp, k))
    V, T = 77, 4
    coef = g.new_coefs(V)
    coef[0:3] = gpu.coeffs("lopass", 0.05, 0.7)[:, None]
    coef[3] = 0.25
    st = g.new_state(V)
    st[0] = np.arange(V)
    w = wl.Workload("noise", g, V, coef, st)
    po, _, ps = port.run(g, V, T, None, st, coef)
    go, _, gs, name = run_gpu(gpu, w, T, None)
    assert name.startswith("fused:"), name
    assert_same_bits(go, po)
    assert_state_equal(gs, ps)


def test_contract_s_scalar_param_input(gpu, port):
    """DSPVector(float) broadcast of a per-voice frequency (examples/audio-and-midi/sine.cpp:33)."""
    g = GraphSpec()
    f = g.param()
    s = g.node("SINE", f)
    lp = g.node("LOPASS", s)
    k = g.param()
    g.output(g.node("MULTIPLY", lp, k))
    V, T = 130, 4
    coef = g.new_coefs(V)
    coef[0] = wl.base_freq(V)
    coef[1:4] = gpu.coeffs("lopass", 0.1, 0.5)[:, None]
    coef[4] = 0.1
    st = g.new_state(V)
    st[0] = wl.SINE_ZERO_PHASE
    w = wl.Workload("contract_s", g, V, coef, st)
    po, pm, ps = port.run(g, V, T, None, st, coef, want_mix=True, mix_mode=1)
    go, gm, gs, name = run_gpu(gpu, w, T, None, want_mix=True)
    assert name.startswith("fused:"), name
    assert_same_bits(go, po)
    assert_same_bits(gm, pm)
    assert_state_equal(gs, ps)


def test_config4_fm3_fdn8(gpu, port):
    V, T = 48, 14
    w = wl.config_4(V)
    inp = w.inputs(T)
    po, pm, ps = port.run(w.spec, V, T, inp, w.state, w.coef, want_mix=True, mix_mode=1)
    go, gm, gs, name = run_gpu(gpu, w, T, inp, want_mix=True, splits=(5, 9))
    assert_same_bits(go, po, "fm3_fdn8 " + name)
    assert_same_bits(gm, pm)
    assert_state_equal(gs, ps)
    assert np.abs(go).max() > 0.1  # the reverb tail is really there


@pytest.mark.parametrize("n_voices,n_blocks,splits", [(50, 20, (7, 13)), (4, 3, None), (130, 6, (1, 5))])
def test_fdn8_fused_and_generic_agree_with_oracle(gpu, port, n_voices, n_blocks, splits):
    """FM3 -> FDN8 on the fused kernel and on the graph interpreter, ragged voice counts,
    delays shorter and longer than a block, several launches (ring write index carries over)."""
    w = wl.config_4(n_voices)
    inp = w.inputs(n_blocks)
    po, pm, ps = port.run(w.spec, n_voices, n_blocks, inp, w.state, w.coef, want_mix=True, mix_mode=1)
    for flags, kind in ((0, "fused:fm3_fdn8"), (gpu.FLAG_FORCE_GENERIC, "generic")):
        go, gm, gs, name = run_gpu(gpu, w, n_blocks, inp, flags=flags, want_mix=True, splits=splits)
        assert name.startswith(kind), name
        assert_same_bits(go, po, kind + " out")
        assert_same_bits(gm, pm, kind + " mix")
        assert_state_equal(gs, ps, kind)


def test_fdn8_on_external_input(gpu, port):
    """INPUT -> FDN8: the reverb fed with external audio (NoiseGen rows, seed = voice index)."""
    g = GraphSpec()
    x = g.input(0)
    fl = g.node("FDN8", x)
    fr = g.node("FDN8_R", fl)
    g.output(fl, fr)
    V, T = 37, 9
    coef = g.new_coefs(V)
    for v in range(V):
        coef[:, v] = gpu.coeffs_fdn8(wl.FDN_TIMES + 32 * (v % 5), wl.FDN_CUTOFFS, np.full(8, 0.6, np.float32))
    rng = np.random.default_rng(5)
    inp = (rng.standard_normal((T, 1, V, 64)) * 0.1).astype(np.float32)
    st = g.new_state(V)
    w = wl.Workload("fdn_in", g, V, coef, st)
    po, _, ps = port.run(g, V, T, inp, st, coef)
    go, _, gs, name = run_gpu(gpu, w, T, inp, splits=(4, 5))
    assert name == "fused:fm3_fdn8", name
    assert_same_bits(go, po)
    assert_state_equal(gs, ps)


def test_config5_chain256(gpu, port):
    V, T = 40, 3
    w = wl.config_5(V, 256)
    po, _, ps = port.run(w.spec, V, T, None, w.state, w.coef)
    go, _, gs, name = run_gpu(gpu, w, T, None)
    assert_same_bits(go, po, "chain256 " + name)
    assert_state_equal(gs, ps)


@pytest.mark.parametrize("n_voices", [300, 30000])
def test_process_call_replayed_from_a_cuda_graph(gpu, port, n_voices):
    """Nothing launch-specific is baked into the kernel arguments (the scheduler words maintain themselves on
    the device), so a process call captured once into a CUDA graph can be replayed: 4 replays == 4 calls.
    300 voices = the two-warp team kernel, 30 000 = the persistent grid with state hopping between warps."""
    import torch
    T, reps = 8, 4
    w = wl.config_a(n_voices)
    inp = w.inputs(T)  # the same planes every replay, state carried
    st, want = w.state, None
    outs = []
    for _ in range(reps):
        want, _, st = port.run(w.spec, n_voices, T, inp, st, w.coef, nthreads=8)
        outs.append(want)
    dev = torch.device("cuda", 0)
    g = gpu.VoiceGraph(w.spec, n_voices)
    try:
        g.set_coefs(w.coef)
        g.set_state(w.state)
        d_in = torch.from_numpy(inp).to(dev)
        d_out = torch.empty((T, 1, n_voices, 64), dtype=torch.float32, device=dev)
        d_mix = torch.zeros((T, 1, 64), dtype=torch.float32, device=dev)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            g.process_device(d_in, d_out, d_mix, T, side.cuda_stream)  # first call: allocations, attributes
            side.synchronize()
            g.set_state(w.state)
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg, stream=side):
                g.process_device(d_in, d_out, d_mix, T, side.cuda_stream)
            g.set_state(w.state)  # capture does not execute
            for r in range(reps):
                cg.replay()
                torch.cuda.synchronize()
                assert_same_bits(d_out.cpu().numpy(), outs[r], "replay %d" % r)
        assert_state_equal(g.get_state(), st)
    finally:
        g.close()


@pytest.mark.parametrize("n_voices", [700, 30000])
def test_asynchronous_mix_reduce(gpu, port, n_voices):
    """mlb_graph_set_mix_async: mix_reduce_kernel leaves the caller's stream (partials double-buffered by call
    parity); five back-to-back calls with their own mix buffers, joined by mix_wait, equal the checker."""
    import torch
    T, calls = 4, 5
    w = wl.config_a(n_voices)
    inp = w.inputs(T * calls)
    po, pm, ps = port.run(w.spec, n_voices, T * calls, inp, w.state, w.coef, want_mix=True, mix_mode=1, nthreads=8)
    dev = torch.device("cuda", 0)
    g = gpu.VoiceGraph(w.spec, n_voices)
    try:
        g.set_coefs(w.coef)
        g.set_state(w.state)
        g.set_mix_async(True)
        d_in = torch.from_numpy(inp).to(dev)
        d_out = torch.empty((T * calls, 1, n_voices, 64), dtype=torch.float32, device=dev)
        d_mix = torch.zeros((T * calls, 1, 64), dtype=torch.float32, device=dev)
        sh = torch.cuda.current_stream().cuda_stream
        torch.cuda.synchronize()
        for c in range(calls):
            g.process_device(d_in[c * T:(c + 1) * T], d_out[c * T:(c + 1) * T], d_mix[c * T:(c + 1) * T], T, sh)
        g.mix_wait(sh)
        torch.cuda.synchronize()
        assert_same_bits(d_out.cpu().numpy(), po, "out")
        assert_same_bits(d_mix.cpu().numpy(), pm, "mix (asynchronous reduce)")
        # the host entry point still returns finished results
        o, m = g.process_host(np.ascontiguousarray(inp[:T]), T, want_out=False, want_mix=True)
        assert np.isfinite(m).all()
    finally:
        g.close()
