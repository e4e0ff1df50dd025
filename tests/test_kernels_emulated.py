"""Kernel logic unit tests on the CPU: the same .cu source compiled for the host (tests/emu) must
reproduce the oracle bit for bit.  This is a test build only -- the product never loads it."""
import numpy as np
import pytest
import torch

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
                                         ("delft", 40), ("melbourne_2000", 120),
                                         ("nine_intersections_smulders", 499), ("45_intersections_smulders", 200)])
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


def _lattice(size, origins, steps=200):
    from pednstream_b200 import Network
    from pednstream_b200.grid import DEFAULT_LINK, grid_adjacency
    params = {"unit_time": 10, "simulation_steps": steps, "default_link": dict(DEFAULT_LINK),
              "demand": {f"origin_{o}": {"peak_lambda": 40, "base_lambda": 25} for o in origins}}
    np.random.seed(3)
    return Network(grid_adjacency(size), params, origin_nodes=list(origins), verbose=False)


def test_user_supplied_turning_fractions(emu_lib):
    """Network.update_turning_fractions_per_node (reference network.py:250-255, usage
    examples/long_corridor.py:111-113): host-owned fractions reach the node kernel (tf_mode 1)."""
    rng = np.random.RandomState(0)

    def fractions(node):
        m = node.source_num
        f = rng.uniform(0.1, 1.0, size=(m, m - 1))
        return (f / f.sum(axis=1, keepdims=True)).reshape(-1)

    a, b = _lattice(3, [0, 8]), _lattice(3, [0, 8])
    for net in (a, b):
        rng = np.random.RandomState(0)
        net.update_turning_fractions_per_node([4, 1], [fractions(net.nodes[4]), fractions(net.nodes[1])])
    o = LtmOracle(a)
    attach(b, emu_lib)
    state = np.random.get_state()
    for t in range(1, 150):
        if t == 70:                                   # change them mid-run as well
            rng = np.random.RandomState(9)
            o.tf[4] = fractions(a.nodes[4])
        o.network_loading(t)
    np.random.set_state(state)
    for t in range(1, 150):
        if t == 70:
            rng = np.random.RandomState(9)
            b.update_turning_fractions_per_node([4], [fractions(b.nodes[4])])
        b.network_loading(t)
    for f in F64_FIELDS[:7] + F32_FIELDS:
        assert np.array_equal(o.h[f], b._store.field(f)), f
    assert float(o.h["cumulative_inflow"][149].sum()) > 1000


def test_device_fault_is_raised_like_the_reference(emu_lib):
    """Negative flows make the reference raise (node.py:219 Warning, link.py:346 ValueError); the
    kernels flag them and the facade raises at the next synchronisation point."""
    net = _lattice(3, [0, 8])
    attach(net, emu_lib)
    net.nodes[0].demand = np.asarray(net.nodes[0].demand, dtype=np.float64)
    net.nodes[0].demand[5:] = -3.0
    with pytest.raises(ValueError, match="negative node flow"):
        for t in range(1, 12):
            net.network_loading(t)
        net.links[(0, 1)].inflow


def test_going_back_in_time_clears_the_flow_rows(emu_lib):
    """The node pass stores nothing for nodes without sending flow and relies on zeroed inflow/outflow
    rows (pns_b200.h row contract).  Moving forward that holds by construction; when a caller repeats
    steps the engine clears the rows first, so values of the earlier pass cannot leak into the new one.
    (A rewind is not a replay -- the running travel-time sum is carried state in the reference too,
    link.py:84,183-186 -- so the check is on the rows themselves.)"""
    steps = 90
    net = make_network("nine_intersections", rng="philox", seed=4)
    eng = attach(net, emu_lib, rng="philox", seed=4)
    for t in range(1, steps + 1):
        net.network_loading(t)
    inflow = eng.history("inflow").clone()
    assert float(inflow[41:steps + 1].abs().sum()) > 0
    # repeat from step 40 with the network emptied of demand: nothing may survive from the first pass
    for n in net.nodes.values():
        if n.demand is not None:
            n.demand[:] = 0
    net.network_loading(40)
    for f in ("inflow", "outflow"):
        h = eng.history(f)
        assert float(h[41:steps + 1].abs().sum()) == 0.0, f          # rows above the repeated step: cleared
        assert torch.equal(h[:40], (inflow if f == "inflow" else h)[:40])
    # row 40 itself holds exactly what the repeated step stored: links next to idle nodes read zeros
    cin = eng.history("cumulative_inflow")
    assert torch.equal(cin[40] - cin[39], eng.history("inflow")[40])
