#!/usr/bin/env python
"""Callback-sized launches: wall time per call of the headline chain at the cadence of an audio callback
(reference: SignalProcessBuffer::process computes ceil(frames / 64) vectors per host callback,
source/app/MLSignalProcessBuffer.cpp:57-78; a 512-frame callback is T = 8 blocks).

For V in {4096, 65536} voices and T in {1, 2, 8} blocks per call:
  host_S   mlb_graph_process_host, contract S (per-voice scalar frequency in, per-voice rows out)
  host_M   mlb_graph_process_host, contract M (scalar frequency in, mix bus out only: 256 B x T back)
  dev_R    mlb_graph_process_device, contract R, buffers resident, launch + synchronize
  graph_R  the same call captured once into a CUDA graph (torch.cuda.CUDAGraph) and replayed
p50 / p99 over `--calls` calls each.  One JSON line per cell -> profiles/latency_r2.jsonl.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pct(xs):
    a = np.sort(np.asarray(xs)) * 1e6
    return {"p50_us": round(float(a[len(a) // 2]), 1), "p99_us": round(float(a[int(len(a) * 0.99)]), 1),
            "min_us": round(float(a[0]), 1)}


def main():
    import torch
    from madronalib_b200 import api, workloads as wl
    from madronalib_b200.graph import GraphSpec, SINE_ZERO_PHASE
    ap = argparse.ArgumentParser()
    ap.add_argument("--calls", type=int, default=400)
    args = ap.parse_args()
    api.init(0)
    dev = torch.device("cuda", 0)
    for V in (4096, 65536):
        w = wl.config_a(V)
        gs = GraphSpec()
        gs.output(gs.node("MULTIPLY", gs.node("LOPASS", gs.node("SINE", gs.param())), gs.param()))
        coef_s = gs.new_coefs(V)
        coef_s[0] = wl.base_freq(V)
        coef_s[1:4] = w.coef[0:3]
        coef_s[4] = w.coef[3]
        st_s = gs.new_state(V)
        st_s[0] = SINE_ZERO_PHASE
        for T in (1, 2, 8):
            rt_us = T * 64 / 48000.0 * 1e6
            # ---- host entry point, contracts S and M ----
            g = api.VoiceGraph(gs, V)
            g.set_coefs(coef_s)
            g.set_state(st_s)
            h_out = torch.empty((T, 1, V, 64), dtype=torch.float32).pin_memory().numpy()
            h_mix = np.empty((T, 1, 64), np.float32)
            for name, want_out in (("host_S", True), ("host_M", False)):
                for _ in range(20):
                    g.process_host(None, T, want_out=want_out, want_mix=True, out=h_out if want_out else None, mix=h_mix)
                ts = []
                for _ in range(args.calls):
                    t0 = time.perf_counter()
                    g.process_host(None, T, want_out=want_out, want_mix=True, out=h_out if want_out else None, mix=h_mix)
                    ts.append(time.perf_counter() - t0)
                print(json.dumps({"cell": name, "voices": V, "blocks": T, "audio_us_per_call": round(rt_us, 1),
                                  "kernel": g.kernel_name, **pct(ts)}), flush=True)
            g.close()
            # ---- device-resident contract R: direct launch vs CUDA-graph replay ----
            g = api.VoiceGraph(w.spec, V)
            g.set_coefs(w.coef)
            g.set_state(w.state)
            d_in = torch.from_numpy(w.inputs(T)).to(dev)
            d_out = torch.empty((T, 1, V, 64), dtype=torch.float32, device=dev)
            d_mix = torch.zeros((T, 1, 64), dtype=torch.float32, device=dev)
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                for _ in range(20):
                    g.process_device(d_in, d_out, d_mix, T, side.cuda_stream)
                side.synchronize()
                ts = []
                for _ in range(args.calls):
                    t0 = time.perf_counter()
                    g.process_device(d_in, d_out, d_mix, T, side.cuda_stream)
                    side.synchronize()
                    ts.append(time.perf_counter() - t0)
                print(json.dumps({"cell": "dev_R", "voices": V, "blocks": T, "audio_us_per_call": round(rt_us, 1),
                                  "kernel": g.kernel_name, **pct(ts)}), flush=True)
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg, stream=side):
                    g.process_device(d_in, d_out, d_mix, T, side.cuda_stream)
                for _ in range(20):
                    cg.replay()
                side.synchronize()
                torch.cuda.synchronize()
                ts = []
                for _ in range(args.calls):
                    t0 = time.perf_counter()
                    cg.replay()
                    torch.cuda.synchronize()
                    ts.append(time.perf_counter() - t0)
                print(json.dumps({"cell": "graph_R", "voices": V, "blocks": T, "audio_us_per_call": round(rt_us, 1),
                                  "kernel": g.kernel_name, **pct(ts)}), flush=True)
            g.close()


if __name__ == "__main__":
    main()
