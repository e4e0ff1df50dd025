"""Parity tests proper: the CUDA path (through the C-ABI) vs the reference fixtures and the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import assert_matches_golden, load_golden, make_network
from oracle import philox as ph
from oracle.ltm_oracle import F32_FIELDS, F64_FIELDS, LtmOracle
from oracle.philox import PhiloxDraws
from pednstream_b200 import _native
from pednstream_b200.engine import Engine
from pednstream_b200.grid import build_grid_plan

pytestmark = pytest.mark.gpu
FIELDS = F64_FIELDS[:7] + F32_FIELDS


def run_numpy_mode(case, steps):
    net = make_network(case, rng="numpy", device="cuda:0")
    for t in range(1, steps + 1):
        net.network_loading(t)
    return net


# same seed as the reference => same trajectory (fp64 counters and fp32 state bit-equal, which is
# inside BASELINE.json's 1e-9 relative tolerance)
@pytest.mark.parametrize("case,steps", [("long_corridor_example", 499), ("long_corridor", 599),
                                         ("nine_intersections", 499), ("45_intersections", 699),
                                         ("butterfly_scA", 599), ("small_network", 499),
                                         ("one_intersection_v0", 599), ("od_flow_example", 499),
                                         ("delft", 120), ("melbourne_2000", 1999),
                                         # randomised scenarios (randomize_network, env_loader.py:160-424)
                                         ("45_intersections_rand7", 400), ("nine_intersections_rand3", 300),
                                         ("delft_rand11", 80),
                                         # smulders fundamental diagram (functions.py:124-128)
                                         ("nine_intersections_smulders", 499), ("45_intersections_smulders", 400)])
def test_cuda_numpy_mode_matches_reference_fixture(case, steps):
    gold = load_golden(case)
    net = run_numpy_mode(case, steps)
    fields = {f: net._store.field(f) for f in FIELDS}
    assert_matches_golden(gold, fields, steps, int(gold["n_links"]))
    net.engine.check_errors()


def test_cuda_within_tolerance_of_oracle_and_invariants():
    """The stated tolerance (1e-9 relative, fp64) checked explicitly against the oracle, plus
    conservation: cum_in - cum_out == pedestrians on the link; node inflow == node outflow."""
    steps = 200
    a = make_network("nine_intersections")
    want = LtmOracle(a).run(steps)
    b = run_numpy_mode("nine_intersections", steps)
    for f in ("cumulative_inflow", "cumulative_outflow", "density", "inflow", "outflow"):
        got = b._store.field(f)
        np.testing.assert_allclose(got, want[f], rtol=1e-9, atol=0)
    L = len(b.links)
    stock = b._store.field("cumulative_inflow")[:, :L] - b._store.field("cumulative_outflow")[:, :L]
    assert np.allclose(stock[: steps + 1], b._store.field("num_pedestrians")[: steps + 1], atol=1e-3)
    for node in b.nodes.values():
        i_side = sum(l.outflow for l in node.incoming_links)
        o_side = sum(l.inflow for l in node.outgoing_links)
        assert np.array_equal(i_side, o_side)


def test_read_surface_after_stepping():
    net = run_numpy_mode("nine_intersections", 50)
    l = net.links[(4, 5)]
    assert l.inflow.shape == (501,) and l.density.dtype == np.float32
    assert l.cumulative_inflow[50] == l.inflow[:51].sum()
    assert l.get_density(50) == (l.num_pedestrians[50] + l.reverse_link.num_pedestrians[50]) / l.area
    tf = net.nodes[4].turning_fractions
    assert tf.shape == (net.nodes[4].edge_num,) and abs(tf.reshape(4, 3).sum(axis=1) - 1).max() < 1e-3
    assert isinstance(l.inflow.tolist(), list)


def test_philox_samplers_match_python_restatement():
    lib = _native.load()
    n = 4096
    rng = np.random.default_rng(5)
    trials = rng.integers(0, 1400, n).astype(np.int32)
    trials[:64] = rng.integers(0, 4, 64)
    p = rng.uniform(0.02, 0.98, n)
    p[100:200] = 0.9
    dt = torch.from_numpy(trials).cuda()
    dp = torch.from_numpy(p).cuda()
    oi = torch.zeros(n, dtype=torch.int32, device="cuda")
    od = torch.zeros(4 * n, dtype=torch.float64, device="cuda")
    seed, t = 0x1234567890ABCDEF, 77
    for kind, site in ((0, 1), (0, 3), (3, 1), (3, 3), (1, 4), (2, 0)):
        rc = lib.pns_rng_selftest(kind, n, C.c_void_p(dt.data_ptr()), C.c_void_p(dp.data_ptr()), seed, t, site,
                                  C.c_void_p(oi.data_ptr()), C.c_void_p(od.data_ptr()), None)
        assert rc == 0
        torch.cuda.synchronize()
        if kind == 0:
            want = [ph.binomial_philox(seed, t, i, 0, site, int(trials[i]), float(p[i])) for i in range(n)]
            assert oi.cpu().numpy().tolist() == want
        elif kind == 3:     # release / blockers draws from the link's shared block
            want = [ph.binomial_u(int(trials[i]), 0.9 if site == 3 else float(p[i]),
                                  ph.link_draws(seed, t, i, 0)[site // 2]) for i in range(n)]
            assert oi.cpu().numpy().tolist() == want
        elif kind == 1:
            want = np.array([ph.normal_quad_philox(seed, t, i, 0, site) for i in range(n)]).reshape(-1)
            assert np.array_equal(od.cpu().numpy(), want)
        else:
            want = np.array([float(ph.det_pow08(np.float32(p[i]))) for i in range(n)])
            assert np.array_equal(od.cpu().numpy()[:n], want)


@pytest.mark.parametrize("case,steps", [("nine_intersections", 120), ("45_intersections", 60),
                                         ("butterfly_scA", 120), ("long_corridor", 150)])
def test_cuda_philox_mode_matches_oracle(case, steps):
    a = make_network(case)
    want = LtmOracle(a, draws=PhiloxDraws(seed=11)).run(steps)
    b = make_network(case, rng="philox", seed=11, device="cuda:0")
    for t in range(1, steps + 1):
        b.network_loading(t)
    for f in FIELDS:
        assert np.array_equal(want[f], b._store.field(f)), f


def _engine_from_network(net, replicas, rng, seed):
    eng = Engine(net.plan, replicas=replicas, rng=rng, seed=seed, device="cuda:0")
    S = net.simulation_steps
    demand = np.zeros((S + 1, max(1, net.plan["n_demand_rows"])))
    for row, node in enumerate(net.plan["demand_nodes"]):
        d = np.asarray(node.demand, dtype=np.float64)
        demand[: len(d), row] = d
    od_w = (np.stack([net.od_manager.od_flows[k] for k in net.plan["od_keys"]], axis=1)
            if net.od_manager is not None else None)
    tf, supplied = net._static_fractions()
    eng.initialise(net._store.gate, net._store.sep_np64, tf, demand, od_w, supplied)
    return eng


def test_batched_replicas_are_independent_and_keyed():
    """R replicas advance in one launch; replica r equals a single-replica oracle with key r."""
    R, steps, seed = 5, 60, 3
    net = make_network("nine_intersections")
    eng = _engine_from_network(net, R, "philox", seed)
    eng.run(1, steps)
    eng.check_errors()
    for r in (0, 3, 4):
        want = LtmOracle(make_network("nine_intersections"), draws=PhiloxDraws(seed=seed, replica=r)).run(steps)
        for f in FIELDS:
            got = eng.history(f)[:, :, r].cpu().numpy()
            assert np.array_equal(want[f], got), (f, r)


def test_table_mode_multi_step_replays_numpy_mode():
    """Record the draws of a numpy-mode run, then replay all steps in one fused call."""
    steps = 150
    a = make_network("45_intersections", rng="numpy", device="cuda:0")
    L = len(a.links)
    rows_b = np.zeros((steps, 3, L), dtype=np.int32)
    rows_n = np.zeros((steps, L), dtype=np.float64)
    rows_e = np.zeros((steps, a.engine.net.n_opts), dtype=np.float64)
    for t in range(1, steps + 1):
        a.network_loading(t)
        host = a.engine._draw_host.numpy()
        rows_b[t - 1] = host[: 3 * L].reshape(3, L)
        rows_n[t - 1] = host[4 * L: 6 * L].view(np.float64)
        rows_e[t - 1] = a.engine._exp_val_host.numpy()         # the logit exponentials, from the host's numpy
    b = make_network("45_intersections")
    eng = _engine_from_network(b, 1, "table", 0)
    eng.set_draw_table(torch.from_numpy(rows_b).cuda(), torch.from_numpy(rows_n).cuda(), torch.from_numpy(rows_e).cuda())
    eng.run(1, steps, _native.RNG_TABLE)
    eng.check_errors()
    for f in FIELDS:
        assert np.array_equal(a._store.field(f)[: steps], eng.history(f)[:steps, :, 0].cpu().numpy()), f


def test_large_grid_invariants():
    """Size-independent properties at a size the oracle cannot reach: 256x256 lattice, 300 steps."""
    size, steps = 256, 300
    plan, gate, tf, demand = build_grid_plan(size, steps + 1, locality_order=True)
    eng = Engine(plan, replicas=1, rng="philox", seed=1, device="cuda:0")
    eng.initialise(gate, None, tf, demand, None)
    eng.run(1, steps)
    eng.check_errors()
    L = plan["n_links"]
    cin = eng.history("cumulative_inflow")[:, :, 0]
    cout = eng.history("cumulative_outflow")[:, :, 0]
    num = eng.history("num_pedestrians")[:, :, 0].double()
    assert torch.allclose((cin - cout)[: steps + 1, :L], num[: steps + 1], atol=1e-2)
    for f in ("inflow", "outflow", "num_pedestrians", "density", "speed", "travel_time"):
        assert (eng.history(f)[: steps + 1] >= 0).all(), f
    # pedestrians are conserved: everything that entered through origins is on a link or has left
    entered = eng.history("cumulative_outflow")[steps, L:, 0].sum()      # virtual in-links feed the net
    left = eng.history("cumulative_inflow")[steps, L:, 0].sum()          # virtual out-links drain it
    on_links = (cin - cout)[steps, :L].sum()
    assert abs(float(entered - left - on_links)) < 1e-6
    assert float(entered) > 1e4
    # flows are integral except where the gate capacity binds (fractional sending flow)
    assert float(eng.history("speed")[steps].max()) <= 1.1 + 0.5


def test_lane_kernel_equals_pair_kernel_on_large_lattice(monkeypatch):
    """The single-replica link kernel (lane per link: shuffles, L2 prefetch, launch order, shortcuts for
    empty links, skipped hand-over stores) against the plain pair-per-thread kernel that the batched and the
    host-emulation paths use, on a lattice with demand high enough to jam the origin links: all 13 series
    bit-equal over 400 steps."""
    size, steps = 192, 400
    plan, gate, tf, demand = build_grid_plan(size, steps + 1, locality_order=True, peak_lambda=90, base_lambda=60)
    a = Engine(plan, replicas=1, rng="philox", seed=9, device="cuda:0")
    a.initialise(gate, None, tf, demand, None)
    a.run(1, steps)
    a.check_errors()
    monkeypatch.setenv("PNS_PAIR_THREADS", "1")
    b = Engine(plan, replicas=1, rng="philox", seed=9, device="cuda:0")
    b.initialise(gate, None, tf, demand, None)
    b.run(1, steps)
    b.check_errors()
    monkeypatch.delenv("PNS_PAIR_THREADS")
    assert float(a.history("num_pedestrians")[steps].max()) > 800          # jammed links exist
    for f in FIELDS:
        assert torch.equal(a.history(f)[: steps + 1], b.history(f)[: steps + 1]), f


def test_streamed_run_equals_resident_run():
    """Engine.run_streamed (per-step H2D demand row + D2H metric inside the native loop) must produce
    the same trajectory as a device-resident run, and the metric must be the pedestrian count."""
    size, steps = 64, 120
    plan, gate, tf, demand = build_grid_plan(size, steps + 1, locality_order=True)
    a = Engine(plan, replicas=1, rng="philox", seed=5, device="cuda:0")
    a.initialise(gate, None, tf, demand, None)
    a.run(1, steps)
    b = Engine(plan, replicas=1, rng="philox", seed=5, device="cuda:0")
    b.initialise(gate, None, tf, np.zeros_like(demand), None)          # device table starts empty
    host_demand = torch.zeros(tuple(b.demand.shape), dtype=torch.float64).pin_memory()
    host_demand[: demand.shape[0], : demand.shape[1]] = torch.from_numpy(demand)
    from pednstream_b200 import _native
    metric_raw = torch.zeros((steps, _native.METRIC_ROW), dtype=torch.float64).pin_memory()
    b.run_streamed(1, steps, host_demand, metric_raw)
    torch.cuda.synchronize()
    b.check_errors()
    for f in FIELDS:
        assert torch.equal(a.history(f)[: steps + 1], b.history(f)[: steps + 1]), f
    metric = torch.from_numpy(b.streamed_metric(metric_raw, steps))
    want = a.history("num_pedestrians")[1: steps + 1, :, 0].double().sum(dim=1).cpu()
    assert torch.equal(metric, want) and float(metric[-1]) > 0


def _lattice(size, origins, steps=200, **kw):
    from pednstream_b200 import Network
    from pednstream_b200.grid import DEFAULT_LINK, grid_adjacency
    params = {"unit_time": 10, "simulation_steps": steps, "default_link": dict(DEFAULT_LINK),
              "demand": {f"origin_{o}": {"peak_lambda": 40, "base_lambda": 25} for o in origins}}
    np.random.seed(3)
    return Network(grid_adjacency(size), params, origin_nodes=list(origins), verbose=False, **kw)


@pytest.mark.parametrize("mode", ["numpy", "philox"])
def test_cuda_unrouted_lattice_matches_oracle(mode):
    """Single-replica lattice through the lane-per-link kernels vs the oracle, with default and with
    user-supplied turning fractions (network.py:250-255), including a mid-run change."""
    steps = 160
    rng = np.random.RandomState(0)

    def fractions(node):
        m = node.source_num
        f = rng.uniform(0.1, 1.0, size=(m, m - 1))
        return (f / f.sum(axis=1, keepdims=True)).reshape(-1)

    a = _lattice(6, [0, 35, 5])
    b = _lattice(6, [0, 35, 5], rng=mode, seed=4, device="cuda:0")
    for net in (a, b):
        rng = np.random.RandomState(0)
        net.update_turning_fractions_per_node([14, 21], [fractions(net.nodes[14]), fractions(net.nodes[21])])
    o = LtmOracle(a, draws=PhiloxDraws(seed=4) if mode == "philox" else None)
    state = np.random.get_state()
    for t in range(1, steps):
        if t == 80:
            rng = np.random.RandomState(9)
            o.tf[14] = fractions(a.nodes[14])
        o.network_loading(t)
    np.random.set_state(state)
    for t in range(1, steps):
        if t == 80:
            rng = np.random.RandomState(9)
            b.update_turning_fractions_per_node([14], [fractions(b.nodes[14])])
        b.network_loading(t)
    for f in FIELDS:
        assert np.array_equal(o.h[f], b._store.field(f)), f
    # counters of the virtual origin/destination links
    for node in b.nodes.values():
        if node.virtual_incoming_link is not None:
            col = node.virtual_incoming_link._col
            assert np.array_equal(o.h["cumulative_outflow"][:, col], node.virtual_incoming_link.cumulative_outflow)
            assert np.array_equal(o.h["cumulative_inflow"][:, col + 1], node.virtual_outgoing_link.cumulative_inflow)
    assert float(o.h["cumulative_inflow"][steps - 1].sum()) > 1000


def test_fused_multi_step_equals_single_steps():
    """run(t0, n) (UPDATE+FLOWS fused across steps, chained launches) vs n calls of run(t, 1)."""
    plan, gate, tf, demand = build_grid_plan(48, 101, locality_order=True)
    a = Engine(plan, replicas=1, rng="philox", seed=9, device="cuda:0")
    a.initialise(gate, None, tf, demand, None)
    a.run(1, 100)
    b = Engine(plan, replicas=1, rng="philox", seed=9, device="cuda:0")
    b.initialise(gate, None, tf, demand, None)
    for t in range(1, 101):
        b.run(t, 1)
    for f in FIELDS:
        assert torch.equal(a.history(f), b.history(f)), f
