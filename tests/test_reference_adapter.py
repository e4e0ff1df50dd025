"""INTEGRATION.md section 2, runnable: the CUDA timestep attached to the reference's *own* Network object
(pednstream_b200.reference_adapter.B200Step) leaves that object in the state its own network_loading would --
every per-link array, host edits between steps included.  Needs a reference tree (/root/reference in the build
container, or the baseline/_ref install); kernels run in the host-emulation build here and on the GPU below."""
import numpy as np
import pytest

from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")

FIELDS = rh.LINK_FIELDS


def _reference_run(dataset, steps, edits):
    np.random.seed(0)
    net, _ = rh.create_network(dataset)
    for t in range(1, steps + 1):
        edits(net, t)
        net.network_loading(t)
    return net


def _adapter_run(dataset, steps, edits, **kw):
    from pednstream_b200.reference_adapter import B200Step
    np.random.seed(0)
    net, _ = rh.create_network(dataset)
    before = np.random.get_state()[2]
    b200 = B200Step(net, **kw).install()
    assert np.random.get_state()[2] == before, "attaching moved numpy's global stream"
    for t in range(1, steps + 1):
        edits(net, t)
        net.network_loading(t)              # now the device step
    b200.facade.engine.check_errors()
    return net


def _edits(net, t):
    """Host mutations between steps, on the reference objects (examples/long_corridor.py:65-66, 124-134)."""
    if (4, 5) in net.links:
        if t == 40:
            net.links[(4, 5)].back_gate_width = 0.8
        if t == 90:
            net.links[(4, 5)].back_gate_width = 3.5
    origin = net.origin_nodes[0]
    if t == 60:
        net.nodes[origin].demand[60:70] = 55


def _compare(a, b, steps):
    assert list(a.links.keys()) == list(b.links.keys())
    for key in a.links:
        for f in FIELDS:
            x, y = np.asarray(getattr(a.links[key], f)), np.asarray(getattr(b.links[key], f))
            upto = steps if f in ("sending_flow", "receiving_flow") else steps + 1
            assert x.dtype == y.dtype and np.array_equal(x[:upto], y[:upto]), (key, f)
    for nid in a.nodes:
        ta, tb = a.nodes[nid].turning_fractions, b.nodes[nid].turning_fractions
        assert (ta is None) == (tb is None) and (ta is None or np.array_equal(ta, tb)), nid


@pytest.mark.parametrize("dataset,steps", [("nine_intersections", 140), ("long_corridor", 160)])
def test_adapter_on_reference_network_emulated(dataset, steps, emu_lib):
    want = _reference_run(dataset, steps, _edits)
    got = _adapter_run(dataset, steps, _edits, _lib=emu_lib, _emulation=True)
    _compare(want, got, steps)


def test_adapter_with_optimal_node_model_emulated(emu_lib):
    """assign_flows_type 'optimal' on the reference's own object: the device solves the node programs itself, so
    the trajectory is not linprog's (tests/test_optimal_node_model.py states what is equal); here: it runs, the
    reference object's arrays are filled, pedestrians are conserved and every node passes on what it receives."""
    from pednstream_b200.reference_adapter import B200Step
    np.random.seed(0)
    net, _ = rh.create_network("nine_intersections", params={"assign_flows_type": "optimal"})
    assert net.assign_flows_type == "optimal"
    b200 = B200Step(net, _lib=emu_lib, _emulation=True).install()
    assert b200.facade.engine.net.n_lp_nodes > 0
    steps = 120
    for t in range(1, steps + 1):
        net.network_loading(t)
    b200.facade.engine.check_errors()
    total = 0.0
    for link in net.links.values():
        stock = np.asarray(link.cumulative_inflow)[: steps + 1] - np.asarray(link.cumulative_outflow)[: steps + 1]
        assert np.allclose(stock, np.asarray(link.num_pedestrians)[: steps + 1], atol=1e-3)
        assert (np.asarray(link.inflow)[: steps + 1] >= 0).all() and (np.asarray(link.outflow)[: steps + 1] >= 0).all()
        total += float(np.asarray(link.cumulative_inflow)[steps])
    assert total > 0
    for node in net.nodes.values():
        taken = sum(np.asarray(l.outflow)[: steps + 1] for l in node.incoming_links)
        given = sum(np.asarray(l.inflow)[: steps + 1] for l in node.outgoing_links)
        assert np.array_equal(taken, given), node.node_id


@pytest.mark.gpu
def test_adapter_on_reference_network_cuda():
    want = _reference_run("nine_intersections", 200, _edits)
    got = _adapter_run("nine_intersections", 200, _edits, device="cuda:0")
    _compare(want, got, 200)
