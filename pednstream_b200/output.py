"""History egress: the simulation directory the reference's `OutputHandler.save_network_state` writes
(handlers/output_handler.py:27-93) -- `link_data.json`, `node_data.json`, `network_params.json` with the same
keys -- so that the reference's loader, visualiser and KPI readers (`OutputHandler.load_simulation`,
rl/rl_utils.py:770-1512) work on a run of this package.  The reference's own handler also runs unchanged on the
facade (tests/test_reference_consumers.py); this writer is for environments that do not have the reference tree
on their path (`PedNetParallelEnv.save`)."""
from __future__ import annotations

import json
import os

import numpy as np

_LINK_SERIES = ("density", "link_flow", "speed", "travel_time", "inflow", "outflow", "num_pedestrians",
                "cumulative_inflow", "cumulative_outflow", "sending_flow", "receiving_flow")


def _plain(x):
    """JSON-encodable form of numpy scalars / arrays / nested containers."""
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, np.generic):
        return x.item()
    if isinstance(x, dict):
        return {str(k): _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple, set)):
        return [_plain(v) for v in x]
    return x


def save_network_state(network, directory: str) -> str:
    """Write the three JSON files of a saved simulation into `directory` (created if missing)."""
    os.makedirs(directory, exist_ok=True)
    gaters = getattr(network, "controller_gaters", set())
    links = {}
    for (u, v), link in network.links.items():
        entry = {name: np.asarray(getattr(link, name)).tolist() for name in _LINK_SERIES}
        entry["parameters"] = {"length": link.length, "width": link.width, "free_flow_speed": link.free_flow_speed,
                               "k_critical": link.k_critical, "k_jam": link.k_jam}
        if u in gaters:                     # gate widths are stored once, on the gater's outgoing links
            entry["back_gate_width"] = np.asarray(link.back_gate_width_data).tolist()
        if getattr(link, "is_separator", False):
            entry["is_separator"] = True
            entry["separator_width"] = np.asarray(link.separator_width_data).tolist()
        links[f"{u}-{v}"] = entry
    nodes = {}
    for node in network.nodes.values():
        nodes[str(node.node_id)] = {"demand": np.asarray(node.demand).tolist() if node.demand is not None else [],
                                    "incoming_links": [l.link_id for l in node.incoming_links],
                                    "outgoing_links": [l.link_id for l in node.outgoing_links]}
    pf = getattr(network, "path_finder", None)
    params = {"simulation_steps": network.simulation_steps, "unit_time": network.unit_time,
              "destination_nodes": network.destination_nodes, "origin_nodes": network.origin_nodes,
              "od_paths": {f"{k[0]}-{k[1]}": v for k, v in pf.od_paths.items()} if pf is not None else {}}
    for name, data in (("link_data.json", links), ("node_data.json", nodes), ("network_params.json", params)):
        with open(os.path.join(directory, name), "w") as fh:
            json.dump(_plain(data), fh)
    return directory
