#!/usr/bin/env python
"""Which stage of the small-bank team pipeline is the slow one?  Times fused chains with cheap / expensive
generators and filters at 4096 voices x 64 blocks (kernel-resident, CUDA events)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from madronalib_b200 import api, workloads as wl
    from madronalib_b200.graph import GraphSpec
    api.init(0)
    dev = torch.device("cuda", 0)
    V, T = int(os.environ.get("PROBE_V", 4096)), 64
    cases = []
    cases.append(("sine_lopass", wl.config_2("lopass", V)))
    cases.append(("phasor_lopass_onepole", wl.config_3(V)))
    cases.append(("sine_lopass_gain(A)", wl.config_a(V)))
    g = GraphSpec()
    g.output(g.node("MULTIPLY", g.node("LOPASS", g.node("NOISE")), g.param()))
    c, s = g.new_coefs(V), g.new_state(V)
    c[0:3] = wl._svf_coefs("lopass", V)
    c[3] = 0.1
    s[0] = np.arange(V)
    cases.append(("noise_lopass_gain", wl.Workload("n", g, V, c, s)))
    g = GraphSpec()
    g.output(g.node("SINE", g.input(0)))
    s = g.new_state(V)
    w = wl.Workload("s", g, V, g.new_coefs(V), s)
    w.inputs = wl.config_a(V).inputs
    cases.append(("sine_only", w))
    g = GraphSpec()
    g.output(g.node("LOPASS", g.input(0)))
    c = g.new_coefs(V)
    c[0:3] = wl._svf_coefs("lopass", V)
    w = wl.Workload("l", g, V, c, g.new_state(V))
    w.inputs = wl.config_a(V).inputs
    cases.append(("lopass_only", w))
    for name, w in cases:
        gr = api.VoiceGraph(w.spec, V)
        gr.set_coefs(w.coef)
        gr.set_state(w.state)
        inp = w.inputs(T)
        d_in = torch.from_numpy(inp).to(dev) if inp is not None else None
        d_out = torch.empty((T, 1, V, 64), dtype=torch.float32, device=dev)
        sh = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            gr.process_device(d_in, d_out, None, T, sh)
        ms = []
        for _ in range(20):
            gr.process_device(d_in, d_out, None, T, sh)
            ms.append(gr.last_kernel_ms())
        k = float(np.mean(ms))
        print(json.dumps({"case": name, "voices": V, "kernel_ms": round(k, 5),
                          "cycles_per_sample_at_1965MHz": round(k * 1e-3 * 1.965e9 / (T * 64), 1), "kernel": gr.kernel_name}),
              flush=True)
        gr.close()


if __name__ == "__main__":
    main()
