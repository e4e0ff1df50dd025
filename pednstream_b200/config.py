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


_MISSING = object()
# where every entry of `params` comes from: (params key, yaml section or None for top level, key, default)
_PARAMS_SCHEMA = (
    ("simulation_steps", "simulation", "simulation_steps", _MISSING),
    ("unit_time", "simulation", "unit_time", _MISSING),
    ("assign_flows_type", "simulation", "assign_flows_type", "classic"),
    ("seed", "simulation", "seed", None),
    ("path_finder", "simulation", "path_finder", dict),
    ("default_link", None, "default_link", _MISSING),
    ("links", None, "links", dict),
    ("demand", None, "demand", dict),
    ("controllers", None, "controllers", dict),
)


def _lookup(doc: dict, section, key, default):
    scope = doc if section is None else doc[section]
    if key in scope:
        return scope[key]
    if default is _MISSING:
        raise KeyError(key)
    return default() if default is dict else default


def load_config(config_path: str) -> dict:
    """Returns {'params', 'origin_nodes', 'destination_nodes'[, 'adjacency_matrix'][, 'od_flows']}."""
    with open(config_path, "r") as fh:
        doc = yaml.safe_load(fh)
    network = doc["network"]
    loaded = {"params": {name: _lookup(doc, section, key, default)
                         for name, section, key, default in _PARAMS_SCHEMA},
              "origin_nodes": network["origin_nodes"],
              "destination_nodes": network.get("destination_nodes", [])}
    adjacency = network.get("adjacency_matrix")
    if adjacency is not None:
        loaded["adjacency_matrix"] = np.array(adjacency)
    weights = doc.get("od_flows")
    if weights is not None:                               # keys "o_d" -> (o, d)
        loaded["od_flows"] = {tuple(map(int, name.split("_"))): w for name, w in weights.items()}
    return loaded


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
