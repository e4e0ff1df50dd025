"""The reference's own downstream consumers run unchanged on the facade (SURVEY.md 8b read surface).

`handlers/output_handler.py` of the reference is loaded from the reference tree by file path (it
is not copied); its `from src.LTM...` imports resolve to this repo's `src/` shims, so the handler
sees `pednstream_b200` objects.  Needs the reference tree -> skipped on the GPU box.
"""
import importlib.util
import json
import os

import numpy as np
import pytest

from conftest import load_golden, make_network
from test_kernels_emulated import attach

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _load_handler():
    spec = importlib.util.spec_from_file_location("_ref_output_handler",
                                                  os.path.join(REF, "handlers", "output_handler.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("case,steps", [("nine_intersections", 499), ("butterfly_scA", 300)])
def test_reference_output_handler_saves_facade_network(case, steps, emu_lib, tmp_path):
    gold = load_golden(case)
    net = make_network(case)
    attach(net, emu_lib)
    for t in range(1, steps + 1):
        net.network_loading(t)
    mod = _load_handler()
    from pednstream_b200.network import Network
    assert mod.Network is Network                       # the shim resolved to the facade class
    h = mod.OutputHandler(base_dir=str(tmp_path), simulation_dir="sim")
    h.save_network_state(net)
    h.save_time_series(net)
    data = mod.OutputHandler.load_simulation(str(tmp_path / "sim"))
    link_data = data["link_data"]
    keys = [tuple(int(x) for x in k) for k in gold["link_keys"]]
    assert set(link_data) == {f"{u}-{v}" for u, v in keys}
    for j, col in enumerate(gold["sample_links"]):
        u, v = keys[int(col)]
        entry = link_data[f"{u}-{v}"]
        for f in ("density", "speed", "travel_time", "inflow", "outflow", "num_pedestrians",
                  "cumulative_inflow", "cumulative_outflow", "link_flow"):
            want = gold["sample_" + f][:steps, j]
            got = np.asarray(entry[f][:steps], dtype=want.dtype)
            assert np.array_equal(got, want), (f, u, v)
        link = net.links[(u, v)]
        assert entry["parameters"]["k_jam"] == link.k_jam
        if entry.get("is_separator"):
            assert len(entry["separator_width"]) == len(link.separator_width_data)
    node_data = data["node_data"]
    for k in gold["demand_nodes"]:
        d = np.asarray(node_data[str(int(k))]["demand"])
        assert np.array_equal(d, gold[f"demand_{int(k)}"])
    assert data["network_params"]["simulation_steps"] == net.simulation_steps
    ts = (tmp_path / "sim" / "time_series.csv").read_text().splitlines()
    assert len(ts) == 1 + len(keys) * net.simulation_steps


def test_env_save_writes_the_reference_layout(emu_lib, tmp_path):
    """PedNetParallelEnv.save (reference rl/pz_pednet_env.py:688-691): the directory the package writes is read by
    the reference's own loader and equals, entry by entry, what the reference's own OutputHandler saves from the
    same network."""
    from pednstream_b200.rl import PedNetParallelEnv
    env = PedNetParallelEnv("long_corridor", obs_mode="option1", seed=3, _lib=emu_lib, _emulation=True)
    env.reset()
    for k in range(60):
        env.step({a: env.action_space(a).sample() for a in env.agents})
    assert env.render() is None                                 # no render mode: nothing to draw
    out = env.save("mine", base_dir=str(tmp_path))
    mod = _load_handler()
    mod.OutputHandler(base_dir=str(tmp_path), simulation_dir="theirs").save_network_state(env.network)
    mine = mod.OutputHandler.load_simulation(out)
    theirs = mod.OutputHandler.load_simulation(str(tmp_path / "theirs"))
    assert mine["network_params"] == theirs["network_params"]
    assert mine["node_data"] == theirs["node_data"]
    assert set(mine["link_data"]) == set(theirs["link_data"])
    for key, entry in theirs["link_data"].items():
        assert mine["link_data"][key] == entry, key
    assert any(e.get("is_separator") for e in mine["link_data"].values())
