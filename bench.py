#!/usr/bin/env python
"""bench.py -- headline benchmark: voice-samples/s of the 65 536-voice
SineGen -> Lopass(SVF) -> gain chain at 48 kHz (BASELINE.json metric, SURVEY.md 8d "Config A").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own SSE path on host cores

One "step" = one pass of the hot path over one batch: T = 64 blocks (4096 samples, 85 ms of
audio) for every voice = ONE fused kernel launch (contract R: per-voice signal-rate frequency
rows in, per-voice output rows out, reference layout [T][V][64] f32) plus the mix-bus
partial sums and their tiny reduce kernel.  Inputs are resident in HBM before the timed
region; in + out = 2.1 GB per step per GPU, far larger than the 126 MB L2, so every step
streams from HBM (no L2 flush needed).

N > 1 (torchrun, one process per GPU): weak scaling -- every GPU owns its own bank of
65 536 voices (voices are independent units: no data-path collective); the only exchange
is the NCCL all-reduce of the [T][1][64] mix bus, issued asynchronously so it overlaps the
next step's kernel.  Time = max over ranks, device-timed with CUDA events.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voice_samples_per_sec"
UNIT = "voice-samples/s"
N_VOICES = 65536
N_BLOCKS = 64
BLOCK = 64


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML during the timed region."""

    def __init__(self, index: int, period_s: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.sm_max, self.power = [], set(), None, []
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self.stop_flag.is_set():
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------
# reference arm: the reference's own SSE implementation on the host cores


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_chain_runner(n_voices: int, n_blocks: int, repeats: int = 1):
    """Returns (run_once() -> seconds, kind, threads).  One call = `repeats` passes over
    n_voices x n_blocks of config A with the compiled reference's own functors (oracle/_ref,
    struct Voice{SineGen; Lopass}) on all host threads; falls back to the plain-C port
    (oracle/_port) only where the reference library was never built."""
    from madronalib_b200 import workloads as wl
    from oracle import bindings

    w = wl.config_a(n_voices)
    inp = w.inputs(n_blocks)  # [T][1][V][64]
    nthreads = host_threads()
    if bindings.ref_available():
        R = bindings.RefOracle()
        coef3 = np.ascontiguousarray(w.coef[0:3])
        gain = np.ascontiguousarray(w.coef[3])
        phase = np.ascontiguousarray(w.state[0]).copy()
        ic = np.ascontiguousarray(w.state[1:3]).view(np.float32).copy()
        x = np.ascontiguousarray(inp[:, 0])

        def run_once():
            _, sec = R.chain_sine_lopass_gain(x, coef3, gain, phase, ic, nthreads, repeats)
            return sec
        return run_once, "reference", nthreads
    if not os.path.exists(bindings.PORT_LIB):
        bindings.build("port")
    P = bindings.PortOracle()

    def run_once():
        t0 = time.perf_counter()
        for _ in range(repeats):
            P.run(w.spec, n_voices, n_blocks, inp, w.state, w.coef, nthreads=nthreads)
        return time.perf_counter() - t0
    return run_once, "port", nthreads


REF_VOICES = 32768  # bounded sample: half the bank (2 x 0.5 GB host buffers), same per-voice work
REF_REPEATS = 4     # passes per step inside one thread launch (amortises thread start-up)


def cpu_baseline(budget_s: float = 12.0):
    """Time the reference chain on a bounded sample of the same workload."""
    V = REF_VOICES
    run_once, kind, nthreads = reference_chain_runner(V, N_BLOCKS, REF_REPEATS)
    run_once()  # warm-up (page faults, thread start)
    secs, t_start = [], time.perf_counter()
    while len(secs) < 3 or (time.perf_counter() - t_start < budget_s and len(secs) < 40):
        secs.append(run_once())
    vs = V * N_BLOCKS * BLOCK * REF_REPEATS
    best = float(np.median(secs))
    return {"value": vs / best, "unit": UNIT, "cores": nthreads, "kind": kind,
            "sample": f"{V} voices x {N_BLOCKS} blocks x {REF_REPEATS} passes per call, {len(secs)} calls "
                      f"of config A (median call {best * 1e3:.1f} ms), std::thread x {nthreads}"}


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    V, reps = REF_VOICES, REF_REPEATS
    run_once, kind, nthreads = reference_chain_runner(V, N_BLOCKS, reps)
    t = run_once()
    # keep the whole run within a few minutes: shrink the per-step sample if needed
    total_steps = args.steps + args.warmup
    while t * total_steps > 150.0 and (reps > 1 or V > 1024):
        if reps > 1:
            reps //= 2
        else:
            V //= 2
        run_once, kind, nthreads = reference_chain_runner(V, N_BLOCKS, reps)
        t = run_once()
    for _ in range(args.warmup):
        run_once()
    secs = [run_once() for _ in range(args.steps)]
    vs = V * N_BLOCKS * BLOCK * reps
    total = float(np.sum(secs))
    value = vs * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config A: SineGen->Lopass(SVF)->gain, contract R, 48 kHz; "
                               f"bounded sample {V} voices x {N_BLOCKS} blocks x {reps} passes per step",
                   "voices": V, "blocks_per_step": N_BLOCKS * reps},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": kind,
                         "sample": f"{V} voices x {N_BLOCKS} blocks x {reps} passes per step, "
                                   f"{args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------
# this repo's arm


def run_cuda_arm(args):
    import torch

    from madronalib_b200 import api, workloads as wl

    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; this repo has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local_rank)
    api.init(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29513")
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))

    V, T = args.voices, args.blocks
    dev = torch.device("cuda", local_rank)
    w = wl.config_a(V)
    graph = api.VoiceGraph(w.spec, V, api.FLAG_FAST if args.fast else api.FLAG_EXACT)
    # multi-GPU: SMs can be kept out of the persistent chain grid for the NCCL kernels of the overlapped
    # mix-bus all-reduce.  Measured at 8 GPUs (profiles/scale8_r1.md): 0 reserved 5.59e12, 4 reserved
    # 5.04e12 voice-samples/s -- the all-reduce already overlaps, so nothing is reserved by default.
    reserved_sms = int(os.environ.get("MLB_BENCH_RESERVED_SMS", "0")) if world > 1 else 0
    graph.reserve_sms(reserved_sms)
    graph.set_coefs(w.coef)
    graph.set_state(w.state)

    # host inputs (pinned: also used by the e2e leg), then resident copy in HBM
    h_in = torch.empty((T, 1, V, BLOCK), dtype=torch.float32).pin_memory()
    w.inputs(T, out=h_in.numpy())
    d_in = h_in.to(dev, non_blocking=True)
    d_out = torch.empty((T, 1, V, BLOCK), dtype=torch.float32, device=dev)
    use_mix = bool(args.mix)
    d_mix = [torch.zeros((T, 1, BLOCK), dtype=torch.float32, device=dev) for _ in range(2)]
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    from madronalib_b200.parallel import MixBusReducer
    reducer = MixBusReducer(dist)

    def step(i: int):
        m = d_mix[i & 1]
        reducer.wait(i)  # the all-reduce issued two steps ago on this buffer
        graph.process_device(d_in, d_out, m if use_mix else None, T, sh)
        if use_mix:
            reducer.submit(i, m)  # NCCL all-reduce of the [T][1][64] mix bus, overlaps step i+1

    def drain():
        reducer.drain()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # untimed warm-up: at least 3 steps; with NCCL at least 20, its first collectives set up channels lazily
    n_warm = max(args.warmup, 3 if world == 1 else 20)
    for i in range(n_warm):
        step(i)
    drain()
    barrier()

    sampler = ClockSampler(physical_gpu_index(local_rank)) if rank == 0 else None
    launches0 = api.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.start()
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    drain()
    e1.record(stream)
    barrier()
    if sampler:
        sampler.stop_flag.set()
        sampler.join()
    launches = api.kernel_launches() - launches0
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    vs_per_step = V * T * BLOCK
    value = vs_per_step * args.steps * world / (ms * 1e-3)

    # ---- roofline of the dominant kernel: per-launch CUDA-event durations (library events on
    # the launching stream), measured live, outside the timed region above ----
    kms = []
    for i in range(min(args.steps, 20)):
        graph.process_device(d_in, d_out, d_mix[0] if use_mix else None, T, sh)
        kms.append(graph.last_kernel_ms())
    kernel_ms = float(np.mean(kms))
    n_groups = (V + 31) // 32
    alg_bytes = (2 * V * T * BLOCK * 4              # freq rows in + output rows out (8 B / voice-sample)
                 + V * (2 * 3 * 4 + 4 * 4)          # state r/w 2 x 12 B + coeffs 12 B + gain 4 B per voice
                 + (T * n_groups * BLOCK * 4 if use_mix else 0))  # mix partials
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel": graph.kernel_name, "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                "bytes_per_voice_sample": alg_bytes / vs_per_step}
    traffic_file = os.path.join(ROOT, "profiles", "traffic_latest.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                roofline["traffic"] = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- e2e: the same step through the reference-facing C-ABI call with HOST buffers:
    # pinned host in -> H2D -> kernel -> D2H -> pinned host out, all inside the timed call ----
    h_out = torch.empty((T, 1, V, BLOCK), dtype=torch.float32).pin_memory()
    h_mix = np.empty((T, 1, BLOCK), np.float32)
    n_e2e = max(1, min(args.steps, args.e2e_steps))
    graph.process_host(h_in.numpy(), T, want_out=True, want_mix=use_mix, out=h_out.numpy(), mix=h_mix)
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        graph.process_host(h_in.numpy(), T, want_out=True, want_mix=use_mix, out=h_out.numpy(),
                           mix=h_mix)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = vs_per_step * n_e2e * world / float(t_e2e.item())
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_in.numel() * 4),
           "d2h_bytes_per_step": int(h_out.numel() * 4 + (h_mix.nbytes if use_mix else 0)),
           "steps": n_e2e, "path": "mlb_graph_process_host (pinned host buffers, contract R)"}

    # ---- other I/O contracts of the same chain through the same host entry point:
    # S = per-voice scalar frequency (DSPVector(float) broadcast, examples/audio-and-midi/sine.cpp:33)
    #     in, per-voice rows out;  M = scalar frequency in, mix bus out only (Synth::processVector).
    # Reported beside the graded contract-R e2e, never as a fraction of the HBM roofline.
    variants = {}
    if not args.no_variants:
        from madronalib_b200.graph import GraphSpec, SINE_ZERO_PHASE
        gs = GraphSpec()
        pf = gs.param()
        ps = gs.node("SINE", pf)
        plp = gs.node("LOPASS", ps)
        pk = gs.param()
        gs.output(gs.node("MULTIPLY", plp, pk))
        coef_s = gs.new_coefs(V)
        coef_s[0] = wl.base_freq(V)
        coef_s[1:4] = w.coef[0:3]
        coef_s[4] = w.coef[3]
        st_s = gs.new_state(V)
        st_s[0] = SINE_ZERO_PHASE
        graph_s = api.VoiceGraph(gs, V, api.FLAG_FAST if args.fast else api.FLAG_EXACT)
        graph_s.set_coefs(coef_s)
        graph_s.set_state(st_s)
        for name, want_out in (("contract_S", True), ("contract_M", False)):
            graph_s.process_host(None, T, want_out=want_out, want_mix=True,
                                 out=h_out.numpy() if want_out else None, mix=h_mix)
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                graph_s.process_host(None, T, want_out=want_out, want_mix=True,
                                     out=h_out.numpy() if want_out else None, mix=h_mix)
            dt = time.perf_counter() - t0
            variants[name] = {"value": vs_per_step * n_e2e * world / dt, "unit": UNIT,
                              "h2d_bytes_per_step": 0,
                              "d2h_bytes_per_step": int((h_out.numel() * 4 if want_out else 0) + h_mix.nbytes),
                              "kernel": graph_s.kernel_name}
        graph_s.close()

    line = None
    if rank == 0:
        cpu = cpu_baseline() if (world == 1 and not args.no_cpu_baseline) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "config A: 65536-voice SineGen->Lopass(SVF)->gain, 48 kHz, contract R "
                            "(per-voice freq rows in, per-voice rows out, [T][V][64] f32), "
                            + ("exact (bit-identical to the reference SSE path)" if not args.fast
                               else "fast (FMA contraction allowed)"),
                "voices_per_gpu": V, "blocks_per_step": T, "samples_per_block": BLOCK,
                "mix_bus": use_mix, "parallelism": f"voices x{world} (weak), mix-bus all-reduce",
                "reserved_sms": reserved_sms,
                "l2": "inputs+outputs 2.1 GB per step >> 126 MB L2 (no flush needed)",
                "kernel": graph.kernel_name,
            },
            "roofline": roofline, "e2e": e2e, "e2e_other_contracts": variants,
            "gpu_launches": int(launches),
            "clocks": sampler.summary() if sampler else None,
            "realtime_x": value / (world * V * 48000.0),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    graph.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--voices", type=int, default=N_VOICES)
    ap.add_argument("--blocks", type=int, default=N_BLOCKS)
    ap.add_argument("--mix", type=int, default=1)
    ap.add_argument("--fast", action="store_true", help="allow FMA contraction (not bit-exact)")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the contract S / M e2e legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_cuda_arm(args)


if __name__ == "__main__":
    sys.exit(main())
