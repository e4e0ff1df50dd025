"""SURVEY.md 8c, parity protocol item 2: the on-device (Philox) draw mode replaces the generator, not the
distributions.  256 trajectories of the reference algorithm driven by numpy's MT19937 stream (the oracle with
np.random.seed(k), same fixed demand) against 256 replicas stepped by the kernels with counter-based draws: the
distributions of trajectory-level statistics must agree (two-sample Kolmogorov-Smirnov + means within combined
standard errors)."""
import numpy as np
import pytest

from conftest import make_network
from oracle.ltm_oracle import LtmOracle

N_SEEDS, STEPS = 256, 130
CHECK_ROWS = (40, 80, 130)


def _statistics(num, cum_out, cum_in, dest_links, probe_link):
    """num/cum_* [rows, L(, R)]: per trajectory (a) pedestrians in the network, (b) arrivals at the destinations,
    (c) cumulative inflow of one interior link -- each at three times."""
    out = []
    for t in CHECK_ROWS:
        out += [num[t].sum(axis=0), cum_out[t][dest_links].sum(axis=0), cum_in[t][probe_link]]
    return np.stack([np.atleast_1d(x) for x in out], axis=0)


def _layout(net):
    L = len(net.links)
    dest = [l.index for (u, v), l in net.links.items() if v in set(net.destination_nodes)]
    return L, dest, L // 2


def _mt_worker(job):
    case, seeds = job
    net = make_network(case)
    L, dest, probe = _layout(net)
    rows = []
    for k in seeds:
        np.random.seed(10_000 + k)
        h = LtmOracle(net).run(STEPS)
        rows.append(_statistics(h["num_pedestrians"][:, :L].astype(np.float64), h["cumulative_outflow"][:, :L],
                                h["cumulative_inflow"][:, :L], dest, probe)[:, 0])
    return rows


def _mt_trajectories(case):
    import multiprocessing as mp
    import os
    workers = min(os.cpu_count() or 1, 16)
    jobs = [(case, list(range(w, N_SEEDS, workers))) for w in range(workers)]
    with mp.get_context("fork").Pool(workers) as pool:
        rows = [r for part in pool.map(_mt_worker, jobs) for r in part]
    net = make_network(case)
    _, dest, probe = _layout(net)
    return np.stack(rows, axis=1), net, dest, probe


def _engine_statistics(net, dest, probe, **engine_kw):
    from pednstream_b200.engine import Engine
    eng = Engine(net.plan, replicas=N_SEEDS, rng="philox", seed=77, **engine_kw)
    S = net.simulation_steps
    demand = np.zeros((S + 1, max(1, net.plan["n_demand_rows"])))
    for row, node in enumerate(net.plan["demand_nodes"]):
        d = np.asarray(node.demand, dtype=np.float64)
        demand[: len(d), row] = d
    od_w = np.stack([net.od_manager.od_flows[k] for k in net.plan["od_keys"]], axis=1)
    tf, supplied = net._static_fractions()
    eng.initialise(net._store.gate, net._store.sep_np64, tf, demand, od_w, supplied)
    eng.run(1, STEPS)
    eng.check_errors()
    L = len(net.links)
    num = eng.history("num_pedestrians")[:, :L].double().cpu().numpy()
    cou = eng.history("cumulative_outflow")[:, :L].cpu().numpy()
    cin = eng.history("cumulative_inflow")[:, :L].cpu().numpy()
    return _statistics(num, cou, cin, dest, probe)


def _assert_same_distribution(a, b):
    from scipy import stats
    assert a.shape == b.shape and a.shape[1] == N_SEEDS
    for i in range(a.shape[0]):
        x, y = a[i], b[i]
        if x.std() == 0 and y.std() == 0:
            assert x[0] == y[0]
            continue
        p = stats.ks_2samp(x, y).pvalue
        se = np.sqrt(x.var() / len(x) + y.var() / len(y))
        assert p > 1e-4, (i, p, x.mean(), y.mean())
        assert abs(x.mean() - y.mean()) < 5 * se + 1e-9, (i, x.mean(), y.mean(), se)
        assert 0.6 < (x.std() + 1e-9) / (y.std() + 1e-9) < 1.6, (i, x.std(), y.std())


@pytest.fixture(scope="module")
def mt_runs():
    return _mt_trajectories("nine_intersections")


def test_philox_trajectories_follow_the_mt_distribution_emulated(mt_runs, emu_lib):
    mt, net, dest, probe = mt_runs
    ph = _engine_statistics(net, dest, probe, lib=emu_lib, emulation=True)
    _assert_same_distribution(mt, ph)


@pytest.mark.gpu
def test_philox_trajectories_follow_the_mt_distribution_cuda(mt_runs):
    mt, net, dest, probe = mt_runs
    ph = _engine_statistics(net, dest, probe, device="cuda:0")
    _assert_same_distribution(mt, ph)
