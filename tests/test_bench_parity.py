"""bench.py's parity block (host logic, no GPU): the replay of the sampled voices on the CPU checker
must accept rows that really are the last step's output and flag a single flipped bit."""
import importlib.util
import os

import numpy as np

from madronalib_b200 import workloads as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_parity_check_accepts_truth_and_flags_a_flipped_bit(port):
    bench = _bench()
    V, T, steps = 300, 3, 4
    w = wl.config_a(V)
    inp = w.inputs(T)
    st = w.state
    out = None
    for _ in range(steps):  # what the GPU does: the same input planes every step, state carried
        out, _, st = port.run(w.spec, V, T, inp, st, w.coef)
    sel = np.unique(np.concatenate([np.linspace(0, V - 1, 64).astype(np.int64), [0, V - 1]]))
    rows = np.ascontiguousarray(out[:, 0][:, sel])
    inp_sel = np.ascontiguousarray(inp[:, :, sel])
    ok = bench.parity_check(w, sel, steps, T, inp_sel, rows, st[:, sel])
    assert ok["mismatches"] == 0 and ok["rows"] == len(sel) and ok["steps_replayed"] == steps
    rows2 = rows.copy()
    rows2.view(np.uint32)[1, 5, 17] ^= 1
    bad = bench.parity_check(w, sel, steps, T, inp_sel, rows2, st[:, sel])
    assert bad["row_mismatches"] == 1 and bad["mismatches"] == 1
    st2 = st[:, sel].copy()
    st2[1, 3] ^= 1
    assert bench.parity_check(w, sel, steps, T, inp_sel, rows, st2)["state_mismatches"] == 1


def test_thread_candidates_and_quota():
    bench = _bench()
    c = bench.thread_candidates()
    assert c and c[-1] == bench.host_threads() and all(1 <= t <= bench.host_threads() for t in c)
    txt, cores = bench.cgroup_cpu_quota()
    assert isinstance(txt, str) and (cores is None or cores > 0)
