"""`Network`: the drop-in facade for the reference's LTM network object.

Construction (nodes, link pairs, virtual O/D links, demand draws, OD weights, k-shortest paths)
is host-side Python and follows the reference's construction order (src/LTM/network.py:56-248)
because node/link ordering and global-RNG consumption are observable.  `network_loading(t)`
(reference network.py:266-287) launches the CUDA timestep through the C-ABI in
`csrc/pns_capi.cu`; there is no CPU path for the physics -- without the native library or a
CUDA device `network_loading` raises.

Read surface kept for the reference's consumers (output_handler, visualizer, RL builders):
`.nodes{id: Node}`, `.links{(u,v): Link}`, per-link numpy series, `.params`, `.simulation_steps`,
`.unit_time`, `.origin_nodes`, `.destination_nodes`, `.path_finder`, `.od_manager`,
`.controller_gaters`, `.controller_links`, `.controller_nodes`, `.pos`.
"""
from __future__ import annotations

import logging
from typing import Callable, List

import numpy as np

from .link import Link, Separator
from .node import Node, OneToOneNode, RegularNode
from .od_manager import DemandGenerator, ODManager
from .path_finder import PathFinder
from .plan import compile_plan
from .state import StateStore


class Network:
    @staticmethod
    def setup_logger(log_level=logging.INFO, log_dir=None):
        logger = logging.getLogger(__name__)
        if not logger.handlers:
            h = logging.StreamHandler()
            h.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
            logger.addHandler(h)
            logger.setLevel(log_level)
        return logger

    def __init__(self, adjacency_matrix, params: dict, origin_nodes: list,
                 destination_nodes: list = [], demand_pattern: List[Callable] = None,
                 od_flows: dict = None, pos: dict = None,
                 log_level: int = logging.INFO, verbose: bool = True,
                 rng: str = "numpy", seed: int = 0, device=None, _lib=None, _emulation: bool = False):
        """rng: 'numpy'  -- in-step draws come from numpy's global legacy RNG in the reference's
                            order (same seed => same trajectory as the reference);
                'philox' -- counter-based on-device sampling keyed (seed, t, link, site)."""
        self.verbose = verbose
        self.logger = self.setup_logger(log_level) if verbose else None
        self.adjacency_matrix = adjacency_matrix
        self.nodes = {}
        self.links = {}
        self.params = params
        self.simulation_steps = params["simulation_steps"]
        self.unit_time = params["unit_time"]
        self.destination_nodes = destination_nodes
        self.origin_nodes = origin_nodes
        self.path_finder = None
        self.od_manager = None
        self.pos = pos
        self.assign_flows_type = params.get("assign_flows_type", "classic")
        if self.assign_flows_type not in ("classic", "optimal"):
            raise ValueError(f"unknown assign_flows_type {self.assign_flows_type!r} (reference node.py:248-300 "
                             "knows 'classic' and 'optimal')")
        self._info(f"Network initialization started, assign flows type: {self.assign_flows_type}")

        self.rng_mode = rng
        self.seed = seed
        self._device = device
        self._lib, self._emulation = _lib, _emulation      # test hooks (tests/emu host build of the kernels)
        self._engine = None
        self._plan = None
        self._fractions_dirty = True
        self._store = StateStore(self.simulation_steps)

        self.demand_generator = DemandGenerator(self.simulation_steps, params,
                                                self.logger if verbose else None)
        for func in demand_pattern or []:
            self.demand_generator.register_pattern(func.__name__, func)
            self._info(f"Custom demand pattern registered: {func.__name__}")

        ctl = params.get("controllers", {})
        self.controller_enabled = ctl.get("enabled", False)
        self.controller_nodes = set(map(int, ctl.get("nodes", set())))
        self.controller_gaters = self.controller_nodes.copy()
        self.controller_links = ctl.get("links", [])
        for pair in self.controller_links:
            a, b = pair.split("-")
            self.controller_nodes.add(int(a))
            self.controller_nodes.add(int(b))
        self._info(f"Controller configuration: enabled: {self.controller_enabled}, "
                   f"nodes: {self.controller_nodes}, links: {self.controller_links}")

        self._virtual_cols = []       # virtual links in creation order (columns assigned at freeze)
        self.init_nodes_and_links()
        self._info(f"Network initialized with {len(self.nodes)} nodes and {len(self.links)} links")

        if destination_nodes:
            self.od_manager = ODManager(self.simulation_steps, logger=self.logger if verbose else None)
            self.od_manager.init_od_flows(origin_nodes, destination_nodes, od_flows)
            self.path_finder = PathFinder(self.links, params=self.params,
                                          controller_nodes=self.controller_nodes,
                                          controller_links=self.controller_links,
                                          logger=self.logger if verbose else None)
            self.path_finder.find_od_paths(od_pairs=self.od_manager.od_flows.keys(), nodes=self.nodes)

    def _info(self, msg):
        if self.logger and self.verbose:
            self.logger.info(msg)

    # ------------------------------------------------------------------ construction
    def _make_node(self, node_id: int) -> Node:
        """Node class from in/out degree and O/D membership (reference network.py:141-167)."""
        n_in = np.sum(self.adjacency_matrix[:, node_id])
        n_out = np.sum(self.adjacency_matrix[node_id, :])
        is_od = node_id in self.origin_nodes or node_id in self.destination_nodes
        if n_in == 2 and n_out == 2:
            node = RegularNode(node_id, self) if is_od else OneToOneNode(node_id, self)
            attach = is_od
        elif n_in == 1 and n_out == 1:
            node, attach = OneToOneNode(node_id, self), True
        else:
            node, attach = RegularNode(node_id, self), is_od
        if attach:
            self._attach_virtual_links(node)
        return node

    def _attach_virtual_links(self, node: Node):
        k = len(self._virtual_cols)
        vin = node._create_virtual_link(self._store, ("v", k), "in", True)
        vout = node._create_virtual_link(self._store, ("v", k + 1), "out", False)
        self._virtual_cols += [vin, vout]
        if node.node_id in self.origin_nodes:
            cfg = self.params.get("demand", {}).get(f"origin_{node.node_id}", {})
            node.demand = self.demand_generator.generate_custom(node.node_id,
                                                                cfg.get("pattern", "gaussian_peaks"))
            self._info(f"Total demand of origin node {node.node_id}: {np.sum(node.demand)}")
        else:
            node.demand = np.zeros(self.simulation_steps)

    def _link_params(self, i: int, j: int) -> dict:
        per_link = self.params.get("links", {})
        base = self.params.get("default_link", {})
        for key in (f"{i}_{j}", f"{j}_{i}"):
            if key in per_link:
                return {**base, **per_link[key]}
        return base

    def init_nodes_and_links(self):
        adj = self.adjacency_matrix
        n = adj.shape[0]
        for i in range(n):
            if i not in self.nodes:
                self.nodes[i] = self._make_node(i)
            node_i = self.nodes[i]
            for j in range(i + 1, n):
                if adj[i, j] != 1:
                    continue
                if j not in self.nodes:
                    self.nodes[j] = self._make_node(j)
                node_j = self.nodes[j]
                kw = self._link_params(i, j)
                if f"{i}-{j}" in self.controller_links or f"{j}-{i}" in self.controller_links:
                    ltype = "separator"
                else:
                    ltype = kw.get("controller_type", "gate")
                if ltype not in ("separator", "gate"):
                    raise ValueError(f"Invalid controller type: {ltype}")
                cls = Separator if ltype == "separator" else Link
                base = len(self.links)
                fwd = cls(self._store, base, f"{i}_{j}", node_i, node_j,
                          self.simulation_steps, self.unit_time, **kw)
                rev = cls(self._store, base + 1, f"{j}_{i}", node_j, node_i,
                          self.simulation_steps, self.unit_time, **kw)
                node_i.outgoing_links.append(fwd)
                node_j.incoming_links.append(fwd)
                node_i.incoming_links.append(rev)
                node_j.outgoing_links.append(rev)
                self.links[(i, j)] = fwd
                self.links[(j, i)] = rev
                fwd.reverse_link, rev.reverse_link = rev, fwd
            node_i.init_node()
        self._freeze_columns()

    def _freeze_columns(self):
        links = list(self.links.values())
        L = len(links)
        for k, v in enumerate(self._virtual_cols):
            v._col = L + k
        for idx, node in enumerate(self.nodes.values()):
            node.index = idx
        self._store.freeze(L, len(self._virtual_cols),
                           tt0=[l.travel_time0 for l in links],
                           window=round(100 / self.unit_time),
                           bgw0=[l._width for l in links])

    # ------------------------------------------------------------------ KPIs
    def link_roles(self):
        """(lk_role[L] int32, any_od_path): bit0 link starts at an origin, bit1 ends at a destination,
        bit2 lies on one of the OD paths (the inputs of the reference's KPI readers,
        rl/rl_utils.py:770-1512)."""
        origins, dests = set(self.origin_nodes), set(self.destination_nodes)
        on_path = set()
        if self.path_finder is not None:
            for paths in self.path_finder.od_paths.values():
                for path in paths:
                    on_path.update(zip(path[:-1], path[1:]))
        role = np.zeros(len(self.links), dtype=np.int32)
        for (u, v), link in self.links.items():
            role[link.index] = (1 if u in origins else 0) | (2 if v in dests else 0) | (4 if (u, v) in on_path else 0)
        return role, bool(on_path)

    def kpis(self, t_last: int = None) -> dict:
        """Episode KPIs computed on the device from the history up to row t_last (default: the last row):
        the quantities and ratios of the reference's rl_utils.compute_* readers."""
        t_last = self.simulation_steps if t_last is None else int(t_last)
        from .kpi import kpi_dict
        role, any_path = self.link_roles()
        eng = self.engine
        rows = self.plan["demand_nodes"]
        if rows:                  # the readers sum each origin's whole demand series, also of a partial episode
            table = np.zeros((self.simulation_steps + 1, len(rows)))
            for k, node in enumerate(rows):
                table[: len(node.demand), k] = node.demand
            eng.set_demand(table)
        raw = eng.kpis(t_last, role, any_path).cpu().numpy()
        return kpi_dict(raw)[0]

    # ------------------------------------------------------------------ fractions
    def update_turning_fractions_per_node(self, node_ids: List[int], new_turning_fractions):
        for i, n in enumerate(node_ids):
            self.nodes[n].update_matrix_A_eq(new_turning_fractions[i])

    def _mark_fractions_dirty(self):
        self._fractions_dirty = True

    def _stepped(self):
        return self._engine is not None and self._engine.t_done > 0

    def _routed_fractions(self, node):
        if self._engine is None or self._engine.t_done == 0:
            return None
        return self._engine.routed_fractions(node.index)

    def _static_fractions(self):
        """(concatenated per-node fractions, per-node flag 'user supplied').  Nodes without
        supplied fractions use uniform 1/(m-1) (network.py:269-271), evaluated on the device."""
        out, supplied = [], []
        for node in self.nodes.values():
            tf = node._tf_static
            supplied.append(tf is not None)
            if tf is None:
                tf = (np.ones(node.edge_num) * (1 / (node.dest_num - 1))) if node.edge_num > 0 else np.zeros(0)
            if len(tf) != node.edge_num:
                raise ValueError(f"node {node.node_id}: expected {node.edge_num} turning fractions")
            out.append(np.asarray(tf, dtype=np.float64))
        return (np.concatenate(out) if out else np.zeros(0)), np.asarray(supplied, dtype=bool)

    # ------------------------------------------------------------------ device runtime
    @property
    def plan(self):
        if self._plan is None:
            self._plan = compile_plan(list(self.nodes.values()), list(self.links.values()),
                                      self.unit_time, self.simulation_steps,
                                      self.path_finder, self.od_manager, node_model=self.assign_flows_type)
        return self._plan

    @property
    def engine(self):
        if self._engine is None:
            from .engine import Engine       # imports torch + the native library; fails loudly
            self._engine = Engine(self.plan, replicas=1, device=self._device,
                                  rng=self.rng_mode, seed=self.seed, lib=self._lib, emulation=self._emulation)
            self._engine.bind_network(self)
            self._store.engine = self._engine
        return self._engine

    def network_loading(self, time_step: int):
        """Advance the whole network to `time_step` (>= 1) on the GPU."""
        self.engine.step_network(self, int(time_step))

    def update_link_states(self, time_step: int):
        raise RuntimeError("link state update is fused into network_loading on the device")
