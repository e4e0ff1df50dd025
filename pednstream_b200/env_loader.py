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
        self._merge_link_config(links_cfg, defaults, link_params_overrides)
        if od_flows:
            self.config["od_flows"] = od_flows
        demand_cfg = params.setdefault("demand", {}) if demand_params_overrides else params.get("demand")
        for key, over in (demand_params_overrides or {}).items():
            demand_cfg.setdefault(key, {}).update(over)

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

    def _merge_link_config(self, links_cfg: dict, defaults: dict, overrides: dict = None):
        """Per-link blocks as create_network leaves them (env_loader.py:93-144), in place: overrides are
        merged into the blocks, then every measured edge (u, v) fixes the length of u_v, and of v_u
        unless that direction already has a block of its own.  Note that u_v and v_u may be the same
        dict object after this (and stay so across calls), exactly as in the reference."""
        for link_id, over in (overrides or {}).items():
            links_cfg.setdefault(link_id, {}).update(over)
        for (u, v), dist in (self.network_data["edge_distances"] or {}).items():
            merged = dict(defaults)
            merged.update(links_cfg.get(f"{u}_{v}", {}))
            merged["length"] = dist
            links_cfg[f"{u}_{v}"] = merged
            links_cfg.setdefault(f"{v}_{u}", merged)
        return links_cfg

    def scenario_link_params(self, link_params_overrides: dict = None) -> dict:
        """Parameters every corridor would get from `create_network(..., link_params_overrides=...)`,
        without building a network and without touching the configuration: {(i, j): kwargs} for i < j
        (both directions of a corridor share them, network.py link construction).  Used to build
        per-replica parameter tables for the batched environment."""
        import copy
        params = self.config["params"]
        defaults = params["default_link"]
        cfg = self._merge_link_config(copy.deepcopy(params.get("links", {})), defaults, link_params_overrides)
        adj = np.asarray(self.network_data["adjacency_matrix"])
        out = {}
        for i, j in zip(*np.nonzero(np.triu(adj == 1, 1))):
            i, j = int(i), int(j)
            block = cfg.get(f"{i}_{j}", cfg.get(f"{j}_{i}"))
            out[(i, j)] = {**defaults, **block} if block is not None else dict(defaults)
        return out

    # ------------------------------------------------------------------ domain randomisation
    # (reference env_loader.py:160-424).  The four generators consume the
    # global numpy stream in the reference's order, each re-seeding it with `seed` when one is given,
    # so the same seed yields the same scenario -- and, after it, the same trajectory.
    def randomize_network(self, yaml_file_path: str, seed: int = None, randomize_params: dict = None,
                          verbose: bool = True, **engine_kw):
        """A perturbed copy of the scenario: OD nodes, link bottlenecks, OD weights, demand patterns
        (env_loader.py:160-181).  Like the reference it needs a network created before (controller
        nodes are read from it).  `verbose`/engine keywords are additions of this implementation."""
        self.generate_random_od_nodes(seed)
        link_params = self.generate_random_link_params(seed)
        od_flows = self.generate_random_od_flows(seed)
        demand_params = self.generate_random_demand_params(seed)
        return self.create_network(yaml_file_path, od_flows=od_flows, link_params_overrides=link_params,
                                   demand_params_overrides=demand_params, verbose=verbose, **engine_kw)

    def generate_random_demand_params(self, seed: int = None) -> dict:
        """Per origin: a pattern, a base rate in [2, 10) and a peak rate in [10, 30), at least 5 above
        the base (env_loader.py:183-221).  The seed travels with the parameters."""
        if seed is not None:
            np.random.seed(seed)
        patterns = ["gaussian_peaks", "constant", "sudden_demand"]
        out = {}
        for origin in self.config.get("origin_nodes", []):
            pattern = np.random.choice(patterns)
            base = np.random.uniform(2.0, 10.0)
            peak = max(np.random.uniform(10.0, 30.0), base + 5)
            out[f"origin_{origin}"] = {"pattern": pattern, "base_lambda": float(base),
                                       "peak_lambda": float(peak), "seed": seed}
        return out

    def generate_random_od_flows(self, seed: int = None) -> dict:
        """One weight in [1, 10) per (origin, destination) pair, constant over the episode
        (env_loader.py:223-258)."""
        if seed is not None:
            np.random.seed(seed)
        steps = self.config["params"]["simulation_steps"]
        out = {}
        for o in self.config.get("origin_nodes", []):
            for d in self.config.get("destination_nodes", []):
                if o != d:
                    out[(o, d)] = np.full(steps + 1, np.random.uniform(1.0, 10.0))
        return out

    def _within_two_hops(self, nodes) -> list:
        """Neighbours and neighbours of neighbours, in the iteration order of a Python set (the
        candidate order matters: np.random.choice indexes into it; env_loader.py:283-298)."""
        adj = self.network_data["adjacency_matrix"]
        near = set()
        for n in nodes:
            near.update(np.where(adj[n, :] == 1)[0].tolist())
        second = set()
        for n in near:
            second.update(np.where(adj[n, :] == 1)[0].tolist())
        near.update(second)
        return list(near)

    def generate_random_od_nodes(self, seed: int = None) -> dict:
        """Add / remove / swap origins and add / remove destinations near the configured ones; controller
        nodes never become origins or destinations (env_loader.py:260-359).  Updates the configuration."""
        if seed is not None:
            np.random.seed(seed)
        rnd = np.random
        controllers = self.network.controller_nodes
        origins = self.config.get("origin_nodes", []).copy()
        dests = self.config.get("destination_nodes", []).copy()

        if rnd.random() < 0.5:                                                   # add an origin
            cand = [n for n in self._within_two_hops(origins) if n not in origins and n not in controllers]
            if cand:
                k = rnd.randint(1, min(2, len(cand) + 1))
                origins.extend(int(x) for x in rnd.choice(cand, k, replace=False))
        if len(origins) > 1 and rnd.random() < 0.5:                              # remove one, keep >= 1
            k = rnd.randint(1, min(2, len(origins)))
            drop = rnd.choice(len(origins), k, replace=False)
            origins = [o for i, o in enumerate(origins) if i not in drop]
        if rnd.random() < 0.5:                                                   # move one to a neighbour
            victim = rnd.choice(origins)
            cand = [n for n in self._within_two_hops([victim]) if n not in origins and n not in controllers]
            if cand:
                origins[origins.index(victim)] = int(rnd.choice(cand))

        if rnd.random() < 0.5:                                                   # add destinations
            cand = [n for n in self._within_two_hops(dests) if n not in dests and n not in controllers]
            if cand:
                k = rnd.randint(1, min(3, len(cand) + 1))
                dests.extend(int(x) for x in rnd.choice(cand, k, replace=False))
        if len(dests) > len(origins) and rnd.random() < 0.5:                     # remove destinations
            removable = [d for d in dests if d not in origins]
            if removable:
                k = rnd.randint(1, min(2, len(removable) + 1))
                gone = [int(x) for x in rnd.choice(removable, k, replace=False)]
                dests = [d for d in dests if d not in gone]

        origins, dests = [int(x) for x in origins], [int(x) for x in dests]
        self.config["origin_nodes"], self.config["destination_nodes"] = origins, dests
        return {"origin_nodes": origins, "destination_nodes": dests}

    def generate_random_link_params(self, seed: int = None) -> dict:
        """Local incidents on 20 % of the corridors: a capacity factor in [0.6, 1.2) on k_critical / k_jam
        and/or a free-flow speed factor in [0.6, 0.9), each with probability 1/2
        (env_loader.py:363-424).  Only the u < v direction of a corridor is listed."""
        if seed is not None:
            np.random.seed(seed)
        rnd = np.random
        fallback = self.data_dir.name if self.data_dir.name != "data" else "delft"
        if not self.network_data:
            self.network_data = self.load_network_data(fallback)
        if not self.config:
            self.config = load_config(os.path.join(self.data_dir, fallback, "sim_params.yaml"))
        if self.network_data.get("edge_distances"):
            corridors = [f"{u}_{v}" for (u, v) in self.network_data["edge_distances"].keys() if u < v]
        else:
            rows, cols = np.where(self.network_data["adjacency_matrix"] == 1)
            corridors = [f"{u}_{v}" for u, v in zip(rows, cols) if u < v]
        params = self.config["params"]
        defaults = params["default_link"]
        out = {}
        n_change = int(len(corridors) * 0.2)
        if n_change > 0:
            for link_id in rnd.choice(corridors, n_change, replace=False):
                have = params["links"].get(link_id, {})
                over = {}
                if rnd.random() < 0.5:
                    f = rnd.uniform(0.6, 1.2)
                    over["k_critical"] = max(0.5, have.get("k_critical", defaults["k_critical"]) * f)
                    over["k_jam"] = max(over["k_critical"] * 2.0, have.get("k_jam", defaults["k_jam"]) * f)
                if rnd.random() < 0.5:
                    over["free_flow_speed"] = have.get("free_flow_speed", defaults["free_flow_speed"]) * rnd.uniform(0.6, 0.9)
                if over:
                    out[link_id] = over
        return out
