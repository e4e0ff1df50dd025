"""Kernel logic unit tests on the CPU: the same .cu source compiled for the host (tests/emu) must
reproduce the oracle bit for bit.  This is a test build only -- the product never loads it."""
import numpy as np
import pytest

from conftest import assert_matches_golden, load_golden, make_network
from oracle.ltm_oracle import F32_FIELDS, F64_FIELDS, LtmOracle
from oracle.philox import PhiloxDraws
from pednstream_b200.engine import Engine


def attach(net, emu_lib, rng="numpy", seed=0):
    eng = Engine(net.plan, replicas=1, rng=rng, seed=seed, lib=emu_lib, emulation=True)
    net._engine = eng
    eng.bind_network(net)
    net._store.engine = eng
    return eng


@pytest.mark.parametrize("case,steps", [("long_corridor", 599), ("nine_intersections", 499),
                                         ("butterfly_scA", 599), ("45_intersections", 250),
                                         ("delft", 40), ("melbourne_2000", 120)])
def test_emulated_kernels_match_reference_fixture(case, steps, emu_lib):
    gold = load_golden(case)
    net = make_network(case)
    attach(net, emu_lib)
    for t in range(1, steps + 1):
        net.network_loading(t)
    fields = {f: net._store.field(f) for f in F64_FIELDS[:7] + F32_FIELDS}
    assert_matches_golden(gold, fields, steps, int(gold["n_links"]))


def test_emulated_philox_mode_matches_oracle(emu_lib):
    steps = 60
    a = make_network("nine_intersections")
    want = LtmOracle(a, draws=PhiloxDraws(seed=11)).run(steps)
    b = make_network("nine_intersections", rng="philox", seed=11)
    attach(b, emu_lib, rng="philox", seed=11)
    for t in range(1, steps + 1):
        b.network_loading(t)
    for f in F64_FIELDS[:7] + F32_FIELDS:
        assert np.array_equal(want[f], b._store.field(f)), f


def test_gate_width_edits_between_steps(emu_lib):
    """Host mutations between network_loading calls (examples/long_corridor.py:124-133 usage)."""
    a = make_network("nine_intersections")
    b = make_network("nine_intersections")
    o = LtmOracle(a)
    attach(b, emu_lib)
    state = np.random.get_state()
    for t in range(1, 120):
        if t == 30:
            o.set_back_gate_width((4, 5), 0.5)
        if t == 60:
            o.set_back_gate_width((4, 5), 3.0)
        o.network_loading(t)
    np.random.set_state(state)
    for t in range(1, 120):
        if t == 30:
            b.links[(4, 5)].back_gate_width = 0.5
        if t == 60:
            b.links[(4, 5)].back_gate_width = 3.0
        b.network_loading(t)
    for f in F64_FIELDS[:7] + F32_FIELDS:
        assert np.array_equal(o.h[f], b._store.field(f)), f
    assert b.links[(4, 5)].back_gate_width_data[45] == 0.5
