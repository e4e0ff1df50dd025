"""Host-side construction, plan compilation and the read surface (CPU, no device)."""
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_network
from pednstream_b200 import Link, NetworkEnvGenerator, OneToOneNode, RegularNode, Separator, load_config

DATASETS = sorted(d for d in os.listdir(os.path.join(ROOT, "data"))
                  if os.path.isdir(os.path.join(ROOT, "data", d)))


def test_all_datasets_present():
    assert len(DATASETS) == 12


@pytest.mark.parametrize("name", DATASETS)
def test_build_and_plan(name):
    np.random.seed(0)
    net = NetworkEnvGenerator().create_network(name, verbose=False)
    p = net.plan
    L = len(net.links)
    assert p["n_links"] == L and L % 2 == 0
    keys = list(net.links.keys())
    for i in range(0, L, 2):                      # forward/reverse pairs are adjacent
        assert keys[i] == keys[i + 1][::-1]
    # slot k of both CSR lists refers to the same neighbour
    for n in net.nodes.values():
        for lin, lout in zip(n.incoming_links, n.outgoing_links):
            if lin.is_virtual:
                assert lout.is_virtual
            else:
                assert lin.reverse_link is lout
    meta = p["nd_meta"]
    assert meta.shape == (len(net.nodes), 4)
    assert p["lk_slots"].shape == (L, 2) and p["nd_stride"] in (4, 8)
    assert int((meta[:, 1] & 0xff).sum()) == len(p["nd_in_col"])
    assert len(p["classes"]) >= 1 and p["lk_class"].max() < len(p["classes"])
    assert p["rt_opt_ptr"][-1] == len(p["rt_opt_link"])
    assert p["rt_term_ptr"][-1] == len(p["rt_term_opt"])
    m = meta[:, 1] & 0xff
    assert len(p["rt_term_ptr"]) - 1 == int(sum(m[n] * (m[n] - 1) for n in p["rt_routed_nodes"]))
    assert ((meta[:, 1] >> 16 == 2) == (p["nd_routed"] >= 0)).all()


def test_long_corridor_node_kinds_and_separators():
    np.random.seed(0)
    net = NetworkEnvGenerator().create_network("long_corridor", verbose=False)
    assert all(isinstance(n, OneToOneNode) for n in net.nodes.values())
    assert isinstance(net.links[(2, 3)], Separator) and isinstance(net.links[(3, 2)], Separator)
    assert type(net.links[(0, 1)]) is Link
    sep = net.links[(2, 3)]
    assert sep.separator_width == 2.0 and sep.front_gate_width == 2.0 and sep.area == 200.0
    assert sep.back_gate_width == 2.0 and net.links[(3, 2)].separator_width == 2.0
    sep.separator_width = 2.5
    assert net.links[(3, 2)].separator_width == 1.5 and net.links[(3, 2)].back_gate_width == 1.5
    assert len(sep.separator_width_data) == 601 and sep.separator_width_data[0] == 2.0


def test_gate_width_coupling_and_initial_series():
    np.random.seed(0)
    net = NetworkEnvGenerator().create_network("nine_intersections", verbose=False)
    l = net.links[(4, 5)]
    l.back_gate_width = 1.25
    assert l.back_gate_width == 1.25 and net.links[(5, 4)].front_gate_width == 1.25
    assert l.front_gate_width == 4
    assert isinstance(net.nodes[4], RegularNode) and isinstance(net.nodes[6], OneToOneNode)
    assert l.sending_flow.shape == (501,) and (l.sending_flow == -1).all()
    assert l.travel_time.dtype == np.float32 and l.travel_time[0] == np.float32(50 / 1.1)
    assert (l.avg_travel_time[:10] == l.travel_time[0]).all() and l.avg_travel_time[10] == 0
    assert l.get_density(0) == 0 and l.inflow.tolist()[:3] == [0.0, 0.0, 0.0]
    assert net.nodes[0].virtual_incoming_link.cumulative_outflow.shape == (501,)


def test_demand_and_setup_rng_stream_match_reference_fixture():
    gold = load_golden("45_intersections")
    net = make_network("45_intersections")
    for n in gold["demand_nodes"]:
        assert np.array_equal(net.nodes[int(n)].demand, gold[f"demand_{int(n)}"])
    assert len(net.path_finder.od_paths[(30, 18)]) == 15          # 10 shortest + 9 detours - 4 duplicates


def test_network_loading_without_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    net = make_network("long_corridor")
    with pytest.raises(RuntimeError, match="CUDA"):
        net.network_loading(1)


def test_missing_scenario_raises():
    with pytest.raises(FileNotFoundError):
        NetworkEnvGenerator().create_network("no_such_scenario")


def test_reference_import_paths_resolve_to_this_package():
    """A script written against the reference (`from src.LTM.network import Network`) runs unchanged."""
    import importlib
    import pednstream_b200
    for mod, names in (("src.LTM.network", ["Network"]), ("src.LTM.link", ["Link", "Separator"]),
                       ("src.utils.env_loader", ["NetworkEnvGenerator"]), ("src.utils.config", ["load_config"])):
        m = importlib.import_module(mod)
        for n in names:
            assert getattr(m, n) is getattr(pednstream_b200, n)
