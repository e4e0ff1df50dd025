"""The sparse lattice constructor emits the same plan as the generic (dense-adjacency) path."""
import numpy as np
import pytest

from pednstream_b200 import Network
from pednstream_b200.grid import DEFAULT_LINK, build_grid_plan, default_origins, grid_adjacency


@pytest.mark.parametrize("size,origins", [(3, [0, 8]), (5, None), (6, [0, 5, 14, 35]), (2, [0])])
def test_grid_plan_equals_generic_plan(size, origins):
    S = 80
    origins = default_origins(size, stride=3) if origins is None else origins
    params = {"unit_time": 10, "simulation_steps": S, "default_link": dict(DEFAULT_LINK),
              "demand": {f"origin_{o}": {"peak_lambda": 50, "base_lambda": 30} for o in origins}}
    np.random.seed(0)
    net = Network(grid_adjacency(size), params, origin_nodes=list(origins), verbose=False)
    want = net.plan
    got, gate, tf, demand = build_grid_plan(size, S, origins=origins, demand_seed=0)
    assert list(net.nodes.keys()) == got["node_order"].tolist()
    for k, v in want.items():
        if isinstance(v, np.ndarray):
            assert v.dtype == got[k].dtype and np.array_equal(v, got[k]), k
    for k in ("n_links", "n_nodes", "n_virtual", "n_demand_rows", "n_edges", "window", "unit_time"):
        assert want[k] == got[k], k
    assert np.array_equal(net._store.gate, gate)
    assert np.array_equal(net._static_fractions()[0], tf)
    assert not net._static_fractions()[1].any()
    for row, node in enumerate(want["demand_nodes"]):
        assert np.array_equal(np.asarray(node.demand, dtype=np.float64)[:S], demand[:, row]), node.node_id


def test_lane_block_order_is_a_permutation_that_starts_next_to_the_origins():
    """The launch schedule of the single-replica link kernel (plan.lane_block_order, pns_net.lane_order)."""
    from pednstream_b200.grid import build_grid_plan
    from pednstream_b200.plan import lane_block_order
    plan, _, _, _ = build_grid_plan(48, 20, locality_order=True)
    L, stride, block = plan["n_links"], plan["nd_stride"], 64
    order = lane_block_order(plan["nd_meta"], plan["nd_in_link"], L, stride, block, hops=4)
    n_blocks = (L + block - 1) // block
    assert order.dtype == np.int32 and sorted(order.tolist()) == list(range(n_blocks))
    # links that leave an origin node sit in the leading group
    meta = np.asarray(plan["nd_meta"])
    slots = np.asarray(plan["nd_in_link"]).reshape(len(meta), stride)
    origin_nodes = np.nonzero(meta[:, 2] >= 0)[0]
    first_len = int(np.nonzero(np.diff(order) < 0)[0][0]) + 1
    leading = set(order[:first_len].tolist())
    for n in origin_nodes:
        for col in slots[n]:
            if 0 <= col < L:
                assert col // block in leading and (col ^ 1) // block in leading
    assert first_len < n_blocks                                  # and the rest follows in index order
    assert np.all(np.diff(order[first_len:]) > 0)
    # nothing to schedule without origins
    meta2 = meta.copy(); meta2[:, 2] = -1
    assert lane_block_order(meta2, plan["nd_in_link"], L, stride, block) is None
