"""Turning fractions on the read surface, step by step, against fixtures recorded from the live reference
(PathFinder.update_turning_fractions + check_fractions, path_finder.py:591-715): routed nodes are refreshed on every
step -- idle ones too -- and in the numpy-compatible mode the logit's exponentials come from the host's numpy, so
the fractions are bit-equal to the reference's."""
import numpy as np
import pytest

from conftest import load_golden
from oracle.gen_golden import TF_CASES


def _network(case, **kw):
    from pednstream_b200 import NetworkEnvGenerator
    c = TF_CASES[case]
    np.random.seed(c["seed"])
    return NetworkEnvGenerator().create_network(c["dataset"], verbose=False, **kw)


def _check(case, steps, **kw):
    gold = load_golden(case)
    net = _network(case, **kw)
    routed = gold["routed_nodes"].tolist()
    assert routed, "fixture has no routed nodes"
    for t in range(1, steps + 1):
        net.network_loading(t)
        for n in routed:
            want = gold[f"tf_{n}"][t - 1]
            got = np.asarray(net.nodes[n].turning_fractions)
            assert np.array_equal(got, want), f"node {n}, step {t}: {got} vs {want}"
    # in the first step nothing has entered a physical link yet: every routed node was idle, and its fractions were
    # refreshed (and compared) all the same
    assert net._store.field("outflow")[1, : len(net.links)].sum() == 0
    net.engine.check_errors()


@pytest.mark.parametrize("case,steps", [("tf_nine_intersections", 220), ("tf_45_intersections", 120)])
def test_turning_fractions_match_reference_emulated(case, steps, emu_lib):
    _check(case, steps, _lib=emu_lib, _emulation=True)


@pytest.mark.gpu
@pytest.mark.parametrize("case,steps", [("tf_nine_intersections", 220), ("tf_45_intersections", 160)])
def test_turning_fractions_match_reference_cuda(case, steps):
    _check(case, steps, device="cuda:0")
