"""`NetworkEnvGenerator`: scenario directory -> `Network` (reference: src/utils/env_loader.py:24-158).

Same call surface: `NetworkEnvGenerator(data_dir="data").create_network(name,
custom_demand_functions=None, od_flows=None, link_params_overrides=None,
demand_params_overrides=None)`; `.config`, `.network`, `.network_data` stay public.  It also
accepts `verbose=` (the reference's RL env passes it although the upstream signature rejects it,
SURVEY Q1) and forwards engine options (`rng`, `seed`, `device`) to `Network`.

Per-episode domain randomisation (`randomize_network`, env_loader.py:160-424) is outside the
accelerated path.
"""
from __future__ import annotations

import json
import os
import pickle
from pathlib import Path
from typing import Callable, List

import numpy as np

from .config import load_config
from .network import Network


class NetworkEnvGenerator:
    def __init__(self, data_dir="data"):
        root = Path(__file__).resolve().parent.parent
        self.data_dir = Path(data_dir) if os.path.isabs(str(data_dir)) else root / data_dir
        self.network = None
        self.network_data = None
        self.config = None

    def load_network_data(self, data_path: str) -> dict:
        folder = os.path.join(self.data_dir, f"{data_path}")
        yaml_path = os.path.join(folder, "sim_params.yaml")
        if not os.path.exists(yaml_path):
            raise FileNotFoundError(f"Network data file not found: {yaml_path}")
        self.config = load_config(yaml_path)

        edge_distances = None
        pkl = os.path.join(folder, "edge_distances.pkl")
        if os.path.exists(pkl):
            with open(pkl, "rb") as fh:
                edge_distances = pickle.load(fh)

        if "adjacency_matrix" in self.config:
            adjacency = self.config["adjacency_matrix"]
        else:
            adjacency = np.load(os.path.join(folder, "adj_matrix.npy"))

        positions = None
        pos_json = os.path.join(folder, "node_positions.json")
        if os.path.exists(pos_json):
            with open(pos_json, "r") as fh:
                positions = {str(k): v for k, v in json.load(fh).items()}
        return {"adjacency_matrix": adjacency, "edge_distances": edge_distances,
                "node_positions": positions}

    def create_network(self, yaml_file_path: str, custom_demand_functions: List[Callable] = None,
                       od_flows: dict = None, link_params_overrides: dict = None,
                       demand_params_overrides: dict = None, verbose: bool = True, **engine_kw):
        if self.network_data is None:
            self.network_data = self.load_network_data(yaml_file_path)
        params = self.config["params"]
        defaults = params["default_link"]

        links_cfg = params.setdefault("links", {})
        for link_id, over in (link_params_overrides or {}).items():
            links_cfg.setdefault(link_id, {}).update(over)
        if od_flows:
            self.config["od_flows"] = od_flows
        demand_cfg = params.setdefault("demand", {}) if demand_params_overrides else params.get("demand")
        for key, over in (demand_params_overrides or {}).items():
            demand_cfg.setdefault(key, {}).update(over)

        # measured edge lengths: the (u, v) entry fixes the length of u_v, and of v_u unless that
        # direction already has its own block (env_loader.py:126-144)
        for (u, v), dist in (self.network_data["edge_distances"] or {}).items():
            merged = dict(defaults)
            merged.update(links_cfg.get(f"{u}_{v}", {}))
            merged["length"] = dist
            links_cfg[f"{u}_{v}"] = merged
            links_cfg.setdefault(f"{v}_{u}", merged)

        self.network = Network(
            adjacency_matrix=self.network_data["adjacency_matrix"],
            params=params,
            origin_nodes=self.config.get("origin_nodes", []),
            destination_nodes=self.config.get("destination_nodes", []),
            demand_pattern=custom_demand_functions,
            od_flows=self.config.get("od_flows", None),
            pos=self.network_data.get("node_positions"),
            verbose=verbose, **engine_kw)
        return self.network

    def randomize_network(self, *a, **k):
        raise NotImplementedError("per-episode domain randomisation is not part of the accelerated "
                                  "path (SURVEY.md section 8f.1)")
