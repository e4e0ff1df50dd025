"""The oracle restatement vs fixtures generated from the live reference (CPU)."""
import numpy as np
import pytest

from conftest import assert_matches_golden, load_golden, make_network
from oracle.ltm_oracle import LtmOracle

# (case, steps simulated here) -- prefixes keep the CPU suite short; the fixture has every row
CASES = [("long_corridor_example", 499), ("long_corridor", 599), ("nine_intersections", 200),
         ("butterfly_scA", 300), ("small_network", 300), ("one_intersection_v0", 300),
         ("od_flow_example", 300), ("45_intersections", 120), ("delft", 12), ("melbourne_2000", 40),
         ("nine_intersections_smulders", 499), ("45_intersections_smulders", 150)]


@pytest.mark.parametrize("case,steps", CASES)
def test_oracle_reproduces_reference_trajectory(case, steps):
    gold = load_golden(case)
    net = make_network(case)
    for n in gold["demand_nodes"]:
        assert np.array_equal(net.nodes[int(n)].demand, gold[f"demand_{int(n)}"]), "demand draws differ"
    assert np.array_equal(np.array(list(net.links.keys())), gold["link_keys"])
    h = LtmOracle(net).run(steps)
    assert_matches_golden(gold, h, steps, int(gold["n_links"]))


def test_oracle_invariants():
    """cum_in - cum_out == pedestrians on the link (to fp32 rounding); node conservation."""
    net = make_network("nine_intersections")
    o = LtmOracle(net)
    h = o.run(150)
    L = o.L
    stock = h["cumulative_inflow"][:151, :L] - h["cumulative_outflow"][:151, :L]
    assert np.allclose(stock, h["num_pedestrians"][:151], rtol=0, atol=1e-3)
    for f in ("inflow", "outflow", "num_pedestrians", "density", "speed"):
        assert (h[f][:151] >= 0).all()
    for node in net.nodes.values():
        inflow_side = sum(h["outflow"][:151, l._col] for l in node.incoming_links)
        outflow_side = sum(h["inflow"][:151, l._col] for l in node.outgoing_links)
        assert np.array_equal(inflow_side, outflow_side)
