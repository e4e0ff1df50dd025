"""YAML scenario loader.

Keeps the reference's input schema verbatim (reference: src/utils/config.py:5-52) so the
shipped `data/<name>/sim_params.yaml` files load unchanged:
`simulation{simulation_steps, unit_time, assign_flows_type, seed, path_finder{...}}`,
`default_link{...}`, `links{"u_v": {...}}`, `demand{origin_k: {...}}`,
`controllers{enabled, nodes, links}`, `network{adjacency_matrix, origin_nodes,
destination_nodes}`, `od_flows{"o_d": w}`.
"""
from __future__ import annotations

import numpy as np
import yaml


def load_config(config_path: str) -> dict:
    """Returns {'params', 'origin_nodes', 'destination_nodes'[, 'adjacency_matrix'][, 'od_flows']}."""
    with open(config_path, "r") as fh:
        raw = yaml.safe_load(fh)

    sim = raw["simulation"]
    net = raw["network"]
    out = {
        "params": {
            "simulation_steps": sim["simulation_steps"],
            "unit_time": sim["unit_time"],
            "assign_flows_type": sim.get("assign_flows_type", "classic"),
            "seed": sim.get("seed", None),
            "path_finder": sim.get("path_finder", {}),
            "default_link": raw["default_link"],
            "links": raw.get("links", {}),
            "demand": raw.get("demand", {}),
            "controllers": raw.get("controllers", {}),
        },
        "origin_nodes": net["origin_nodes"],
        "destination_nodes": net.get("destination_nodes", []),
    }
    if "adjacency_matrix" in net:
        out["adjacency_matrix"] = np.array(net["adjacency_matrix"])
    if "od_flows" in raw:
        out["od_flows"] = {tuple(int(x) for x in key.split("_")): w
                           for key, w in raw["od_flows"].items()}
    return out


_REQUIRED = {
    "network": ("origin_nodes",),
    "simulation": ("simulation_steps", "unit_time"),
    "default_link": ("length", "width", "free_flow_speed", "k_critical", "k_jam"),
}


def validate_config(config: dict) -> None:
    """Raises ValueError when a required section/field is missing (reference: config.py:54-77)."""
    for section, fields in _REQUIRED.items():
        if section not in config:
            raise ValueError(f"Missing required section: {section}")
        for field in fields:
            if field not in config[section]:
                raise ValueError(f"Missing required field: {field} in section {section}")
