"""Episode KPIs on the device (C-ABI pns_kpi) against the reference's own readers
(rl/rl_utils.py:770-1512) run on the files the reference's own OutputHandler saves from the facade
network.  Needs the reference tree for the comparison; the invariants run anywhere."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import make_network
from pednstream_b200 import _native
from pednstream_b200.kpi import kpi_dict
from test_kernels_emulated import attach

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _reference_kpis(net, tmp_path):
    handler = _load("handlers/output_handler.py", "_ref_output_handler").OutputHandler(base_dir=str(tmp_path),
                                                                                      simulation_dir="sim")
    handler.save_network_state(net)
    ru = _load("rl/rl_utils.py", "_ref_rl_utils")
    d = str(tmp_path / "sim")
    out = {}
    for fn in ("compute_network_throughput", "compute_served_trips_rate", "compute_average_travel_time_spent",
               "compute_total_network_delay", "compute_network_congestion_metric", "compute_network_travel_time"):
        out[fn] = getattr(ru, fn)(simulation_dir=d)
    return out


@needs_ref
@pytest.mark.parametrize("case", ["nine_intersections", "butterfly_scA", "long_corridor"])
def test_device_kpis_match_reference_readers(case, emu_lib, tmp_path):
    net = make_network(case)
    attach(net, emu_lib)
    S = net.simulation_steps
    for t in range(1, S + 1):                 # a full episode as the environment runs it: row S is written
        net.network_loading(t)
    got = net.kpis()
    ref = _reference_kpis(net, tmp_path)
    exact = [("compute_network_throughput", "completed_demand", "completed_demand"),
             ("compute_network_throughput", "total_demand", "total_demand"),
             ("compute_network_throughput", "throughput", "throughput"),
             ("compute_served_trips_rate", "total_inflow", "total_inflow"),
             ("compute_served_trips_rate", "total_outflow", "total_outflow"),
             ("compute_served_trips_rate", "served_trips_rate", "served_trips_rate"),
             ("compute_average_travel_time_spent", "total_person_time", "total_person_time"),
             ("compute_average_travel_time_spent", "total_trips", "total_trips"),
             ("compute_average_travel_time_spent", "avg_travel_time_spent", "avg_travel_time_spent"),
             ("compute_total_network_delay", "total_person_time", "total_person_time_moving"),
             ("compute_network_congestion_metric", "total_area_time", "total_area_time"),
             ("compute_network_congestion_metric", "congestion_fraction", "congestion_fraction")]
    for fn, rk, gk in exact:                  # sums of integers (or counts): independent of the summation order
        assert got[gk] == pytest.approx(ref[fn][rk], rel=1e-13, abs=0), (fn, rk, got[gk], ref[fn][rk])
    close = [("compute_total_network_delay", "total_delay", "total_delay"),
             ("compute_total_network_delay", "delay_intensity", "delay_intensity"),
             ("compute_network_congestion_metric", "congestion_time", "congestion_time"),
             ("compute_network_congestion_metric", "avg_congestion_density", "avg_congestion_density"),
             ("compute_network_travel_time", "avg_travel_time", "avg_travel_time")]
    for fn, rk, gk in close:                  # real-valued sums: per-link partial sums vs one running sum
        assert got[gk] == pytest.approx(ref[fn][rk], rel=1e-10, abs=1e-12), (fn, rk, got[gk], ref[fn][rk])
    assert got["total_demand"] > 0 and got["total_person_time"] > 0


def test_kpi_invariants(emu_lib):
    net = make_network("nine_intersections")
    attach(net, emu_lib)
    for t in range(1, 201):
        net.network_loading(t)
    k = net.kpis(t_last=200)
    num = net._store.field("num_pedestrians")[:201, :len(net.links)]
    assert k["total_person_time"] == float(num.astype(np.float64).sum() * net.unit_time)
    assert 0.0 <= k["delay_intensity"] <= 1.0 and 0.0 <= k["congestion_fraction"] <= 1.0
    cin = net._store.field("cumulative_inflow")[200, :len(net.links)]
    starts_at_origin = [l.index for (u, v), l in net.links.items() if u in set(net.origin_nodes)]
    assert k["total_inflow"] == float(cin[starts_at_origin].sum())
    assert k["total_demand"] == float(sum(net.nodes[o].demand.sum() for o in net.origin_nodes))
    raw = np.zeros((2, len(_native.KPI_NAMES)))
    assert kpi_dict(raw)[1]["throughput"] == 0.0


@pytest.mark.gpu
def test_batched_kpis_cuda_match_facade():
    """Per-replica KPIs of the batched environment equal the single-network KPIs of the same trajectory."""
    from pednstream_b200.rl import BatchedPedNetEnv
    env = BatchedPedNetEnv("nine_intersections", replicas=33, obs_mode="option3", seed=3, device="cuda:0")
    a = torch.full((33, env.n_act), 2.5, dtype=torch.float32, device=env.device)
    for _ in range(120):
        env.step(a)
    raw = env.kpis().cpu().numpy()
    assert raw.shape == (33, len(_native.KPI_NAMES))
    k = kpi_dict(raw)
    num = env.engine.history("num_pedestrians")[:121].double().cpu().numpy()        # [t, L, R]
    for r in (0, 16, 32):
        assert k[r]["total_person_time"] == float(num[:, :, r].sum() * env.network.unit_time)
        assert k[r]["total_outflow"] >= 0 and k[r]["total_inflow"] > 0
    assert len({round(x["total_delay"], 6) for x in k}) > 1          # replicas differ
