#!/usr/bin/env python
"""Kernel-resident timing of everything except the headline: BASELINE.json configurations 2-5, the
SURVEY 8(f) workloads (config 6 = Aaltoverb x V, the EventsToSignals::Voice bank) and the elementwise maps.

    python tools/bench_configs.py [--steps K] [--only 2,3,4,5,6,voices,map] [--generic] [--cpu]

Prints one JSON line per configuration: kernel name, mean CUDA-event kernel time, algorithmic
bytes per launch (SURVEY.md 8d / DESIGN.md), achieved GB/s and fraction of the measured HBM peak.
--cpu also times the reference (oracle/_ref) on the box's host cores for config 6 and the Voice bank.
Not the headline benchmark (that is bench.py / config A).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from madronalib_b200 import api, workloads as wl
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--only", default="2,3,swept,4,5,6,voices,map")
    ap.add_argument("--generic", action="store_true", help="force the graph interpreter kernel")
    ap.add_argument("--cpu", action="store_true",
                    help="also time the reference (oracle/_ref, all host threads) on config 6 and the Voice bank")
    args = ap.parse_args()
    only = set(args.only.split(","))
    api.init(0)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    dev = torch.device("cuda", 0)
    cfgs = []
    if "2" in only:
        for kind in ("lopass", "hipass", "bell"):
            cfgs.append(("config2_" + kind, wl.config_2(kind, 4096), 64, lambda V, T: 8.0 * V * T * 64))
    if "3" in only:
        cfgs.append(("config3", wl.config_3(65536), 64, lambda V, T: 8.0 * V * T * 64))
    if "swept" in only:
        # the swept headline chain: SineGen -> Lopass(coefficient ROWS g0, g1, g2) -> gain; per voice-sample
        # 4 B freq + 3 x 4 B coefficient rows in + 4 B out = 20 B (MLDSPFilters.h:136-152 after makeCoeffsVec)
        cfgs.append(("configA_swept", wl.swept_filter_case("sine_lopass_v_gain", 65536, 16), 16,
                     lambda V, T: 20.0 * V * T * 64))
    if "4" in only:
        # ring 8 x (256 B r + 256 B w) + freq row in + 2 rows out per voice-block = 4864 B
        cfgs.append(("config4", wl.config_4(16384), 16, lambda V, T: 4864.0 * V * T))
    if "5" in only:
        cfgs.append(("config5", wl.config_5(1024, 256), 16, lambda V, T: 256.0 * V * T))
    if "6" in only:
        # Aaltoverb x V.  Per reverb-block: 12 rings x (256 B write + 256 B read: the two
        # taps of a PitchbendableDelay sit within a sample of each other) + 10 vy1 rows r/w + 2 feedback
        # rows r/w + 2 glide rows read + 2 rows in + 2 rows out = 13824 B
        for V6 in (4096, 16384):
            cfgs.append(("config6_aaltoverb", wl.config_6(V6), 16, lambda V, T: 13824.0 * V * T))
    if "map" in only:
        # K3: stateless elementwise ops, n_rows x 64 elements resident in HBM
        n_rows = 1 << 20  # 64 Mi elements = 256 MB per operand (> L2)
        x1 = torch.rand((n_rows, 64), dtype=torch.float32, device=dev) * 4 - 2
        x2 = torch.rand((n_rows, 64), dtype=torch.float32, device=dev) + 0.5
        x3 = torch.rand((n_rows, 64), dtype=torch.float32, device=dev)
        y = torch.empty_like(x1)
        sh = torch.cuda.current_stream().cuda_stream
        for op, nin in (("multiply", 2), ("sin", 1), ("exp", 1), ("log", 1), ("sin_approx", 1),
                        ("pow", 2), ("lerp", 3), ("clamp", 3)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                api.map_device(op, x1, x2 if nin > 1 else None, x3 if nin > 2 else None, y, n_rows, sh)
            e0.record()
            for _ in range(args.steps):
                api.map_device(op, x1, x2 if nin > 1 else None, x3 if nin > 2 else None, y, n_rows, sh)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            b = (nin + 1) * n_rows * 64 * 4.0
            print(json.dumps({"config": "map_" + op, "kernel": "map_kernel", "rows": n_rows, "kernel_ms": ms,
                              "algorithmic_bytes": b, "achieved_gbs": b / (ms * 1e-3) / 1e9,
                              "frac_of_measured_hbm_peak": b / (ms * 1e-3) / 1e9 / peak}), flush=True)
        del x1, x2, x3, y
    if "voices" in only:
        # K7: EventsToSignals::Voice x V.  Per voice-block: 72 B record in, selected rows out.
        V7, T7 = 65536, 64
        # a busy performance: a voice gets note events in 10 % of its vectors, controller moves in 5 %
        ev = wl.voice_events(256, T7, seed=2, density=0.10, ctl=0.05)
        ev = np.ascontiguousarray(np.tile(ev, (1, V7 // 256)))
        prm = wl.voice_bank_params(V7)
        d_ev = torch.from_numpy(ev.view(np.uint8).reshape(T7, V7, 72)).to(dev)
        d_rows = torch.empty((T7, 8, V7, 64), dtype=torch.float32, device=dev)
        sh = torch.cuda.current_stream().cuda_stream
        for mask, label in ((0x03, "pitch+gate"), (0xFF, "all 8 rows")):
            vb = api.VoiceBank(48000.0, *prm)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(2):
                vb.process_device(d_ev, d_rows, T7, mask, sh)
            e0.record()
            for _ in range(args.steps):
                vb.process_device(d_ev, d_rows, T7, mask, sh)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            b = (72.0 + 256.0 * bin(mask).count("1")) * V7 * T7
            print(json.dumps({"config": "voices_" + label, "kernel": "voice_bank_kernel", "voices": V7, "blocks": T7,
                              "kernel_ms": ms, "voice_samples_per_s": V7 * T7 * 64 / (ms * 1e-3),
                              "algorithmic_bytes": b, "achieved_gbs": b / (ms * 1e-3) / 1e9,
                              "frac_of_measured_hbm_peak": b / (ms * 1e-3) / 1e9 / peak}), flush=True)
            vb.close()
        del d_ev, d_rows
    if args.cpu:
        # the reference itself on this box's host cores (oracle/_ref; checker code, used here only as a baseline)
        import time
        from oracle import bindings as ob
        cores = os.cpu_count() or 1
        if "6" in only:
            R = ob.RefOracle()
            w = wl.config_6(1024)
            inp = w.inputs(16)
            best = 1e30
            for _ in range(3):
                t0 = time.perf_counter()
                R.run(w.spec, w.n_voices, 16, inp, w.state, w.coef, want_out=True, nthreads=cores)
                best = min(best, time.perf_counter() - t0)
            print(json.dumps({"config": "config6_aaltoverb_reference_cpu", "threads": cores, "voices": 1024, "blocks": 16,
                              "seconds": best, "voice_samples_per_s": 1024 * 16 * 64 / best,
                              "note": "reference functors behind the graph interpreter of oracle/ref/mlref.cpp, "
                                      "includes building the per-voice functor objects"}), flush=True)
        if "voices" in only:
            B = ob.ref_voice_bank()
            Vc, Tc = 8192, 64
            ev = np.ascontiguousarray(np.tile(wl.voice_events(256, Tc, seed=2, density=0.10, ctl=0.05), (1, Vc // 256)))
            out, sec = B.run(48000.0, *wl.voice_bank_params(Vc), ev, nthreads=cores)
            print(json.dumps({"config": "voices_reference_cpu", "threads": cores, "voices": Vc, "blocks": Tc,
                              "seconds": sec, "voice_samples_per_s": Vc * Tc * 64 / sec,
                              "note": "the reference's own EventsToSignals::Voice (oracle/_ref/libmle2s.so)"}), flush=True)
    for name, w, T, alg in cfgs:
        V = w.n_voices
        g = api.VoiceGraph(w.spec, V, api.FLAG_FORCE_GENERIC if args.generic else 0)
        g.set_coefs(w.coef)
        g.set_state(w.state)
        inp = w.inputs(T)
        d_in = torch.from_numpy(inp).to(dev) if inp is not None else None
        d_out = torch.empty((T, w.spec.n_out, V, 64), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        sh = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            g.process_device(d_in, d_out, None, T, sh)
        ms = []
        for _ in range(args.steps):
            g.process_device(d_in, d_out, None, T, sh)
            ms.append(g.last_kernel_ms())
        kms = float(np.mean(ms))
        b = alg(V, T)
        print(json.dumps({"config": name, "kernel": g.kernel_name, "voices": V, "blocks": T,
                          "kernel_ms": kms, "voice_samples_per_s": V * T * 64 / (kms * 1e-3),
                          "algorithmic_bytes": b, "achieved_gbs": b / (kms * 1e-3) / 1e9,
                          "frac_of_measured_hbm_peak": b / (kms * 1e-3) / 1e9 / peak,
                          "delay_memory_mb": g.delay_bytes / 1e6}), flush=True)
        g.close()
        del d_in, d_out


if __name__ == "__main__":
    main()
