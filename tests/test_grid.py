"""The sparse lattice constructor emits the same plan as the generic (dense-adjacency) path."""
import numpy as np
import pytest

from pednstream_b200 import Network
from pednstream_b200.grid import DEFAULT_LINK, build_grid_plan, default_origins, grid_adjacency


@pytest.mark.parametrize("size,origins,node_model", [(3, [0, 8], "classic"), (5, None, "classic"),
                                                     (6, [0, 5, 14, 35], "classic"), (2, [0], "classic"),
                                                     (5, [0, 12, 24], "optimal")])
def test_grid_plan_equals_generic_plan(size, origins, node_model):
    S = 80
    origins = default_origins(size, stride=3) if origins is None else origins
    params = {"unit_time": 10, "simulation_steps": S, "default_link": dict(DEFAULT_LINK),
              "assign_flows_type": node_model,
              "demand": {f"origin_{o}": {"peak_lambda": 50, "base_lambda": 30} for o in origins}}
    np.random.seed(0)
    net = Network(grid_adjacency(size), params, origin_nodes=list(origins), verbose=False)
    want = net.plan
    got, gate, tf, demand = build_grid_plan(size, S, origins=origins, demand_seed=0, node_model=node_model)
    assert (((np.asarray(got["nd_meta"])[:, 1] >> 8) & 0xff) == 2).any() == (node_model == "optimal")
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


def _routed_pair(size, S, origins, dests, seed):
    """(generic Network, sparse routed plan tuple) of the same routed lattice, same position of numpy's stream."""
    from pednstream_b200.grid import build_routed_grid_plan
    params = {"unit_time": 10, "simulation_steps": S, "default_link": dict(DEFAULT_LINK),
              "demand": {f"origin_{o}": {"peak_lambda": 50, "base_lambda": 30} for o in origins}}
    np.random.seed(seed)
    net = Network(grid_adjacency(size), params, origin_nodes=list(origins), destination_nodes=list(dests), verbose=False)
    np.random.seed(seed)
    sparse = build_routed_grid_plan(size, S, origins, dests, demand_seed=None)
    return net, sparse


@pytest.mark.parametrize("size,origins,dests", [(8, [0, 7], [63]), (12, [0, 5, 77, 143], [11, 130]),
                                                (16, [3, 100, 255], [240, 15])])
def test_routed_grid_plan_equals_generic_plan(size, origins, dests):
    """BASELINE config 4b's constructor: k-shortest paths and turn structures of a routed lattice built without the
    dense adjacency matrix or per-link objects give the plan the generic constructor compiles -- every table."""
    net, (got, gate, tf, demand, od_w) = _routed_pair(size, 60, origins, dests, seed=4)
    want = net.plan
    assert len(want["rt_routed_nodes"]) > 0 and want["od_keys"] == got["od_keys"]
    for k, v in want.items():
        if isinstance(v, np.ndarray):
            if v.dtype.names:
                assert v.tobytes() == got[k].tobytes(), k
            else:
                assert v.dtype == got[k].dtype and np.array_equal(v, got[k]), k
    for k in ("n_links", "n_nodes", "n_virtual", "n_demand_rows", "n_edges", "n_od", "window", "unit_time"):
        assert want[k] == got[k], k
    assert {k: [list(map(int, p)) for p in v] for k, v in net.path_finder.od_paths.items()} == \
           {k: [list(map(int, p)) for p in v] for k, v in got["od_paths"].items()}
    want_od = np.stack([net.od_manager.od_flows[k] for k in want["od_keys"]], axis=1)
    assert np.array_equal(want_od, od_w)
    for row, node in enumerate(want["demand_nodes"]):
        assert np.array_equal(np.asarray(node.demand, dtype=np.float64)[:60], demand[:, row]), node.node_id


def _routed_lattice_vs_oracle(size, steps, origins, dests, **engine_kw):
    from oracle.ltm_oracle import F32_FIELDS, F64_FIELDS, LtmOracle
    from oracle.philox import PhiloxDraws
    from pednstream_b200.engine import Engine
    net, (plan, gate, tf, demand, od_w) = _routed_pair(size, steps + 20, origins, dests, seed=9)
    want = LtmOracle(net, draws=PhiloxDraws(seed=21)).run(steps)
    eng = Engine(plan, replicas=1, rng="philox", seed=21, **engine_kw)
    eng.initialise(gate, None, tf, demand, od_w)
    eng.run(1, steps)
    eng.check_errors()
    L = plan["n_links"]
    for f in F64_FIELDS[:7] + F32_FIELDS:
        got = eng.history(f)[: steps + 1, :, 0].cpu().numpy()
        upto = steps if f in ("sending_flow", "receiving_flow") else steps + 1
        assert np.array_equal(want[f][:upto], got[:upto]), f
    assert want["cumulative_outflow"][steps, :L].sum() > 0


def test_routed_lattice_sparse_plan_matches_oracle_emulated(emu_lib):
    _routed_lattice_vs_oracle(10, 60, [0, 9], [95, 90], lib=emu_lib, emulation=True)


@pytest.mark.gpu
def test_routed_lattice_sparse_plan_matches_oracle_cuda():
    """20x20 routed lattice from the sparse constructor on the GPU (route choice riding in the single-replica link
    kernel + routed node kernel) against the oracle stepping the generic constructor's network."""
    _routed_lattice_vs_oracle(20, 70, [5, 12, 7 * 20 + 0], [6 * 20 + 14, 11 * 20 + 19], device="cuda:0")
