"""The graph interpreter's host-side planner (build_generic in mlb200.cu, reached without a device through
mlb_graph_plan): which pipeline stage a node goes to and how many shared-memory row slots a program needs.  The rules
checked here are the ones the device relies on: a one-block feedback loop, a paired second row and a functor with its
MLB_AGAIN calls never straddle a stage cut; the row pool fits in 227 KB; the plans agree with what the GPU runs printed
(kernel names in profiles/)."""
import numpy as np
import pytest

from madronalib_b200 import api, workloads as wl
from madronalib_b200.graph import OP_NAME, graph_fdn

STAGE_COUNTS = [None, 1, 2, 3, 5, 8, 64]


def _plans(spec, V, monkeypatch):
    for s in STAGE_COUNTS:
        if s is None:
            monkeypatch.delenv("MLB_STAGES", raising=False)
        else:
            monkeypatch.setenv("MLB_STAGES", str(s))
        yield s, api.plan(spec, V)


def _check_rules(g, stage):
    for i in range(g.n_nodes):
        name = OP_NAME[g.ops[i]]
        t = g.again_target(i)
        if t >= 0:
            assert stage[i] == stage[t], "AGAIN node %d and its functor %d are in different stages" % (i, t)
        if name == "FEEDBACK_WRITE":
            assert stage[i] == stage[g.iargs[i]], "feedback loop of node %d is cut" % i
        if name in ("HALFBAND_UP_2", "FDN8_R"):
            assert stage[i] == stage[g.ins[i][0]], "second row %d is cut from its producer" % i
        for src in g.ins[i]:
            if src >= 0 and stage[src] >= 0 and stage[i] >= 0:
                assert stage[src] <= stage[i], "node %d reads a later stage" % i


def test_plans_match_what_the_gpu_ran():
    """kernel names recorded on the B200: config 5 'generic[19 stages, 3 rows]', config 6 at 70 voices
    'generic[6 stages, 11 rows]' (profiles/configs_r2.jsonl, the smoke() line)."""
    assert api.plan(wl.config_5(1024, 256).spec, 1024)[1:] == (19, 3)
    assert api.plan(wl.config_6(70).spec, 70)[1:] == (6, 11)
    assert api.plan(wl.config_6(4096).spec, 4096)[1:] == (2, 11)
    assert api.plan(wl.config_6(16384).spec, 16384)[1:] == (1, 11)


@pytest.mark.parametrize("name", wl.FUNCTOR_CASES + wl.AGAIN_CASES + ("aaltoverb", "chain64"))
def test_stage_cuts_respect_the_rules(monkeypatch, name):
    w = wl.config_6(70) if name == "aaltoverb" else wl.config_5(40, 64) if name == "chain64" else wl.functor_case(name, 40)
    for s, (stage, n_stages, rows) in _plans(w.spec, w.n_voices, monkeypatch):
        _check_rules(w.spec, stage)
        assert 1 <= n_stages <= (s or 64) and rows * 8704 <= 232448


@pytest.mark.parametrize("seed", list(range(8)) + [217, 226, 238] + list(range(300, 316)))
def test_random_graphs_with_and_without_again_nodes(monkeypatch, seed):
    again = 0.5 if seed >= 200 else 0.0
    w = wl.random_graph_workload(seed, 41, 28, hw_approx=False, again_prob=again)
    n_again = sum(w.spec.again_target(i) >= 0 for i in range(w.spec.n_nodes))
    assert (n_again > 0) == (again > 0) or seed >= 300
    for s, (stage, n_stages, rows) in _plans(w.spec, 41, monkeypatch):
        _check_rules(w.spec, stage)


def test_a_functor_and_its_further_calls_pin_the_cut(monkeypatch):
    """upsample2x_osc: the SINE / LOPASS pairs span most of the program -- however many stages are asked for, everything
    between a functor and its AGAIN call is one stage."""
    g = wl.functor_case("upsample2x_osc", 40).spec
    for s, (stage, n_stages, rows) in _plans(g, 40, monkeypatch):
        first = min(g.again_target(i) for i in range(g.n_nodes) if g.again_target(i) >= 0)
        last = max(i for i in range(g.n_nodes) if g.again_target(i) >= 0)
        assert len({stage[i] for i in range(first, last + 1) if stage[i] >= 0}) == 1
        assert n_stages <= 3


@pytest.mark.parametrize("size,rows_expected", [(4, 8), (6, 10), (16, 20)])
def test_fdn_written_out_fits_the_row_pool(monkeypatch, size, rows_expected):
    """FDN<SIZE> as IntegerDelay / OnePole / feedback nodes: every line's feedback loop overlaps the others, so the
    whole network is one stage whatever is asked for; SIZE + 4 live rows."""
    g, _ = graph_fdn(size)
    for s, (stage, n_stages, rows) in _plans(g, 45, monkeypatch):
        _check_rules(g, stage)
        assert n_stages == 1 and rows == rows_expected


def test_a_network_too_wide_for_shared_memory_is_refused():
    g, _ = graph_fdn(24)  # 28 live rows x 8.5 KB > 227 KB
    with pytest.raises(api.MlbError) as e:
        api.plan(g, 45)
    assert e.value.code == 5 and "live rows" in str(e.value)


def test_traced_bodies_plan(monkeypatch):
    from tests.test_trace import traced
    for case in ("kitchen", "upsample", "fdn", "rows", "rest", "oversample", "shelf", "chain"):
        g, _, _ = traced(case)
        for s, (stage, n_stages, rows) in _plans(g, 36, monkeypatch):
            _check_rules(g, stage)
