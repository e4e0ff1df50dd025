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
