"""Print the graph interpreter's host-side plan (mlb_graph_plan; no GPU needed) for a named workload:
   python tools/plan.py config5 | config6 | functor:<name> | fdn:<size> | random:<seed> [--voices V] [--stages S]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madronalib_b200 import api, workloads as wl  # noqa: E402
from madronalib_b200.graph import OP_NAME  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--voices", type=int, default=1024)
    ap.add_argument("--stages", type=int, default=0, help="force the stage count (MLB_STAGES)")
    a = ap.parse_args()
    if a.stages:
        os.environ["MLB_STAGES"] = str(a.stages)
    kind, _, arg = a.what.partition(":")
    V = a.voices
    if kind == "config5":
        spec = wl.config_5(V, 256).spec
    elif kind == "config6":
        spec = wl.config_6(V).spec
    elif kind == "functor":
        spec = wl.functor_case(arg, V).spec
    elif kind == "fdn":
        spec = wl.fdn_case(int(arg), V)[0].spec
    elif kind == "random":
        spec = wl.random_graph_workload(int(arg), V, 28, hw_approx=False).spec
    else:
        ap.error("unknown workload")
    stage, n_stages, rows = api.plan(spec, V)
    print("%d nodes, %d voices -> %d stage(s), %d row slots (%.1f KB of shared memory per CTA)"
          % (spec.n_nodes, V, n_stages, rows, rows * 8704 / 1024))
    for s in range(n_stages):
        members = [i for i in range(spec.n_nodes) if stage[i] == s]
        print("  stage %2d: %3d nodes  %s" % (s, len(members), " ".join(OP_NAME[spec.ops[i]].lower() for i in members[:12])
                                             + (" ..." if len(members) > 12 else "")))


if __name__ == "__main__":
    main()
