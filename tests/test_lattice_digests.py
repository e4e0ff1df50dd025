"""The single-replica lattice kernel path (k_link_lane with one parameter class, host-resolved lag rows, launch
order, quiet-link shortcut) against committed digests of the oracle run in on-device (Philox) draw mode on jammed
32x32 and 64x64 lattices (oracle/gen_golden.py LATTICE_CASES; the Python restatement of the device samplers)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, row_digests
from oracle.gen_golden import LATTICE_CASES
from oracle.ltm_oracle import F32_FIELDS, F64_FIELDS

FIELDS = F64_FIELDS[:7] + F32_FIELDS


def _run_engine(name, steps, locality_order, **engine_kw):
    import torch
    from pednstream_b200.engine import Engine
    from pednstream_b200.grid import build_grid_plan, default_origins
    case = LATTICE_CASES[name]
    gold = load_golden(name)
    S = int(gold["sim_steps"])
    plan, gate, tf, _ = build_grid_plan(case["size"], S, origins=default_origins(case["size"], case["stride"]),
                                        locality_order=locality_order)
    assert plan["n_links"] == int(gold["n_links"])
    # the fixture's demand, row = the plan's demand row of that node
    order = plan["node_order"]
    rows = {int(order[i]): int(plan["nd_meta"][i, 2]) for i in range(len(order)) if plan["nd_meta"][i, 2] >= 0}
    demand = np.zeros((S, plan["n_demand_rows"]))
    for col, node in enumerate(gold["demand_node_ids"].tolist()):
        demand[:, rows[node]] = gold["demand"][:, col]
    eng = Engine(plan, replicas=1, rng="philox", seed=int(gold["seed"]), **engine_kw)
    eng.initialise(gate, None, tf, demand, None)
    eng.run(1, steps)
    eng.check_errors()
    return gold, eng


def _compare(gold, eng, steps):
    L = int(gold["n_links"])
    for f in FIELDS:
        got = eng.history(f)[: steps + 1, :L, 0].cpu().numpy()
        upto = steps if f in ("sending_flow", "receiving_flow") else steps + 1
        bad = np.nonzero(row_digests(np.ascontiguousarray(got))[:upto] != gold["rows_" + f][:upto])[0]
        assert not len(bad), f"{f}: first differing row {int(bad[0])}"


def test_lattice_fixture_is_jammed():
    for name in LATTICE_CASES:
        if os.path.exists(os.path.join(GOLDEN, name + ".npz")):
            g = load_golden(name)
            assert g["max_pedestrians"] >= 1100 and g["jammed_link_steps"] > 500 and g["steps_run"] >= 300


def test_pair_kernel_matches_lattice_digest_emulated(emu_lib):
    """Host build of the pair-per-thread kernel, first 60 steps of the 32x32 case (the CPU suite stays short)."""
    gold, eng = _run_engine("lattice32_philox", 60, False, lib=emu_lib, emulation=True)
    _compare(gold, eng, 60)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(LATTICE_CASES))
@pytest.mark.parametrize("locality_order", [True, False])
def test_lane_kernel_matches_lattice_digest_cuda(name, locality_order):
    if not os.path.exists(os.path.join(GOLDEN, name + ".npz")):
        pytest.skip("fixture not generated")
    steps = int(load_golden(name)["steps_run"])
    gold, eng = _run_engine(name, steps, locality_order, device="cuda:0")
    _compare(gold, eng, steps)
    assert float(eng.history("num_pedestrians")[:steps].max()) >= 1100      # jammed links: long blockers draws
