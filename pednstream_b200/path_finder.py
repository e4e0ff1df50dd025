"""Route-choice *setup* (host side, runs once per network).

K-shortest simple paths per OD pair, controller-node detours and the static per-node route
structures stay in Python (BASELINE.json north_star: "K-shortest paths stay as host-side
setup"); the per-step logit over these structures runs on the GPU (`csrc/pns_kernels.cu`,
route kernel).  What is built here, and in which order, follows the reference
(src/LTM/path_finder.py:146-546) because the orders are observable in results:

* `od_paths[(o, d)]`            networkx `shortest_simple_paths` on a DiGraph whose edges are
                                inserted in `network.links` order (path ties depend on it);
* `node.turns_distances[od][up][down]`  shortest remaining distance through each turn, dict
                                insertion order = order of first appearance while walking paths;
* `node.up_od_probs[up][od]`    the ODs registered at each upstream (normalisation order);
* `node.ods_in_turns[turn]`     a Python *set* of ODs: the kernel accumulates
                                P(down|up,od)·P(od|up) in this set's iteration order.

`export_route_plan` flattens them into index arrays for the device.
"""
from __future__ import annotations

from collections import defaultdict

import networkx as nx
import numpy as np


def enumerate_shortest_simple_paths(graph, origin, dest, max_paths=None):
    """First `max_paths` loop-free paths by increasing weight ([] if networkx refuses the query)."""
    try:
        it = nx.shortest_simple_paths(graph, origin, dest, weight="weight")
    except Exception:
        return []
    found = []
    for p in it:
        found.append(p)
        if max_paths is not None and len(found) >= max_paths:
            break
    return found


class PathFinder:
    def __init__(self, links, params=None, controller_nodes=None, controller_links=None, logger=None):
        self.links = links
        self.logger = logger
        self.od_paths = {}
        self.nodes_in_paths = set()
        self.node_to_od_pairs = {}
        self.node_turn_probs = {}
        self._initialized = False

        self.graph = nx.DiGraph()
        for (u, v), link in links.items():
            self.graph.add_edge(u, v, weight=link.length, num_pedestrians=0.0)

        cfg = params.get("path_finder", {}) if params else {}
        self.temp = cfg.get("temp", 0.1)
        self.alpha = cfg.get("alpha", 1.0)
        self.beta = cfg.get("beta", 0.05)
        self.omega = cfg.get("omega", 0.05)
        self.std_dev = cfg.get("std_dev", 0)
        # one global-RNG normal is consumed here even when std_dev == 0 (path_finder.py:163)
        self.epsilon = np.random.normal(0, self.std_dev)
        self.k_paths = cfg.get("k_paths", 3)
        self.verbose = cfg.get("verbose", True)

        self.controller_nodes = controller_nodes
        self.controller_links = controller_links
        self.controllers_enabled = bool(controller_nodes or controller_links)
        self.detour_exploration_mode = "penalize"
        self.detour_penalty_factor = 2
        self.max_detour_paths = 3

    def _log(self, msg):
        if self.logger and self.verbose:
            self.logger.info(msg)

    def is_controller_node(self, node_id):
        return bool(self.controllers_enabled and node_id in self.controller_nodes)

    # ------------------------------------------------------------------ path enumeration
    def _register_path_nodes(self, path, od_pair):
        for n in path:
            self.nodes_in_paths.add(n)
            self.node_to_od_pairs.setdefault(n, set()).add(od_pair)

    def find_od_paths(self, od_pairs, nodes):
        for origin, dest in od_pairs:
            try:
                paths = enumerate_shortest_simple_paths(self.graph, origin, dest, max_paths=self.k_paths)
                self.od_paths[(origin, dest)] = paths
                for p in paths:
                    self._register_path_nodes(p, (origin, dest))
            except nx.NetworkXNoPath:
                self._log(f"No path found between {origin} and {dest}")
                self.od_paths[(origin, dest)] = []

        if not self._initialized and self.controllers_enabled:
            for cnode in self.controller_nodes:
                for od_pair in self.node_to_od_pairs[cnode]:
                    before = len(self.od_paths[od_pair])
                    self.expand_controller_paths(nodes[cnode], od_pair)
                    self._log(f"Controller node {cnode}: Added {len(self.od_paths[od_pair]) - before} "
                              f"detour path(s) for OD {od_pair}")
        self._drop_duplicate_paths()
        for node_id in self.nodes_in_paths:
            if nodes[node_id].source_num > 2:
                self.calculate_turn_probabilities(nodes[node_id])
        self._initialized = True

    def _drop_duplicate_paths(self):
        for od_pair, paths in self.od_paths.items():
            as_tuples = [tuple(int(n) for n in p) for p in (paths or [])]
            distinct = set(as_tuples)
            if len(distinct) != len(as_tuples):
                self._log(f"Warning: duplicate paths detected for OD {od_pair}")
                self.od_paths[od_pair] = [list(p) for p in distinct]

    def calculate_path_distance(self, path, start_idx=0):
        total = 0
        for a, b in zip(path[start_idx:-1], path[start_idx + 1:]):
            edge = self.graph.edges[(a, b)]
            if edge:
                total += edge["weight"]
        return total

    def expand_controller_paths(self, current_node, od_pair):
        """Give a controller node alternatives: for every path through it, try the neighbours the
        path does not use and splice in up to `max_detour_paths` loop-free continuations found on
        a graph whose already-used edges are penalised by 1..detour_penalty_factor with distance
        to the destination (path_finder.py:304-458)."""
        here = current_node.node_id
        origin, dest = od_pair
        paths = self.od_paths[od_pair]

        neighbours = set()
        for link in current_node.outgoing_links:
            if link.end_node is not None:
                neighbours.add(link.end_node.node_id)

        biased = self.graph.copy()
        used = {}
        for p in paths:
            for a, b in zip(p[:-1], p[1:]):
                if (a, b) not in used:
                    try:
                        used[(a, b)] = nx.shortest_path_length(self.graph, b, dest, weight="weight")
                    except nx.NetworkXNoPath:
                        used[(a, b)] = 0
        if self.detour_exploration_mode == "remove":
            biased.remove_edges_from([e for e in used if biased.has_edge(*e)])
        elif used:
            far = max(used.values())
            for (a, b), d in used.items():
                if not biased.has_edge(a, b):
                    continue
                if far > 0:
                    factor = 1.0 + (self.detour_penalty_factor - 1.0) * (d / far)
                else:
                    factor = self.detour_penalty_factor
                biased[a][b]["weight"] = biased[a][b].get("weight", 1) * factor

        additions = []
        for path in paths:
            if here not in path:
                continue
            k = path.index(here)
            if here == dest:
                continue
            up = -1 if (here == origin or k == 0) else path[k - 1]
            on_path_down = path[k + 1] if k < len(path) - 1 else None
            for nb in neighbours:
                if nb == on_path_down or nb == up:
                    continue
                if nb in set(path[:k]):
                    continue
                try:
                    tails = enumerate_shortest_simple_paths(biased, nb, dest, max_paths=self.max_detour_paths)
                    if not tails:
                        continue
                    visited = set(path[:k + 1])
                    for tail in tails:
                        if set(tail[1:]) & visited:
                            continue
                        candidate = path[:k + 1] + tail
                        if tuple(candidate) not in set(tuple(p) for p in self.od_paths[od_pair]):
                            additions.append(candidate)
                except Exception:
                    continue

        if additions:
            self.od_paths[od_pair].extend(additions)
            for p in additions:
                self._register_path_nodes(p, od_pair)
        return additions

    # ------------------------------------------------------------------ per-node structures
    def calculate_turn_probabilities(self, node):
        here = node.node_id
        for od_pair in self.node_to_od_pairs.get(here, set()):
            origin, dest = od_pair
            best = {}                      # turn -> shortest remaining distance
            turn = None
            for path in self.od_paths[od_pair]:
                if here not in path:
                    continue
                k = path.index(here)
                if here == origin:
                    turn = (-1, path[k + 1])
                elif here == dest:
                    turn = (path[k - 1], -1)
                elif k < len(path) - 1:
                    turn = (path[k - 1], path[k + 1])
                remaining = self.calculate_path_distance(path, start_idx=k)
                if turn not in best or remaining < best[turn]:
                    best[turn] = remaining
                    if not self._initialized:
                        node.ods_in_turns.setdefault(turn, set()).add(od_pair)

            if best:
                if not hasattr(node, "node_turn_probs"):
                    node.node_turn_probs = {}
                if not hasattr(node, "turns_distances"):
                    node.turns_distances = {}
                if not hasattr(node, "up_od_probs"):
                    node.up_od_probs = defaultdict(lambda: defaultdict(int))
                node.turns_distances[od_pair] = {}
                for (up, down), dist in best.items():
                    node.turns_distances[od_pair].setdefault(up, {})[down] = dist
                    node.up_od_probs[up][od_pair] = 0
                node.node_turn_probs.setdefault(od_pair, {})

    # ------------------------------------------------------------------ device plan export
    def export_route_plan(self, nodes_in_order, od_index, link_index):
        """Flatten the per-node route structures into CSR arrays.

        nodes_in_order: list of Node in network.nodes order; od_index: {(o,d): column in the OD
        weight table}; link_index: {(u,v): physical link index}.
        Returns a dict of int32/float64 numpy arrays (see include/pns_b200.h, `pns_net`).
        """
        grp_node, grp_up, grp_od, grp_has_virtual, opt_ptr = [], [], [], [], [0]
        opt_link, opt_slot, opt_dist = [], [], []
        row_ptr, row_od = [0], []           # per (routed node, up slot): registered ODs
        node_row0 = []                      # first row index of each routed node
        term_ptr, term_opt, term_row_entry = [0], [], []
        routed_nodes, node_edge0 = [], []
        edge_cursor = 0

        for node in nodes_in_order:
            if not (node.node_id in self.nodes_in_paths and node.source_num > 2):
                continue
            m = node.source_num
            ups = [l.start_node.node_id if l.start_node is not None else -1 for l in node.incoming_links]
            downs = [l.end_node.node_id if l.end_node is not None else -1 for l in node.outgoing_links]
            tdist = getattr(node, "turns_distances", {})
            upod = getattr(node, "up_od_probs", {})
            routed_nodes.append(node.index)
            node_edge0.append(edge_cursor)

            group_of = {}                   # (od, up) -> (group id, {down: option index})
            for od, by_up in tdist.items():
                for up, by_down in by_up.items():
                    if not by_down:
                        continue
                    g = len(grp_node)
                    grp_node.append(node.index)
                    grp_up.append(ups.index(up))
                    grp_od.append(od_index[od])
                    has_virtual = 0
                    where = {}
                    for down, dist in by_down.items():
                        where[down] = len(opt_link)
                        key = (node.node_id, down)
                        if key in self.links:
                            opt_link.append(link_index[key])
                        else:
                            opt_link.append(-1)
                            has_virtual = 1
                        opt_slot.append(downs.index(down))
                        opt_dist.append(float(dist))
                    grp_has_virtual.append(has_virtual)
                    opt_ptr.append(len(opt_link))
                    group_of[(od, up)] = (g, where)

            # P(od|up) rows, one per upstream slot (empty when nothing is registered there)
            node_row0.append(len(row_ptr) - 1)
            entry_of = {}
            for i, up in enumerate(ups):
                if up in upod:
                    for od in upod[up]:
                        entry_of[(up, od)] = len(row_od)
                        row_od.append(od_index[od])
                row_ptr.append(len(row_od))

            # accumulation terms per turn, in the iteration order of the Python set
            for up in ups:
                for down in downs:
                    if up == down:
                        continue
                    for od in node.ods_in_turns.get((up, down), set()):
                        g, where = group_of[(od, up)]
                        if down in where and (up, od) in entry_of:
                            term_opt.append(where[down])
                            term_row_entry.append(entry_of[(up, od)])
                    term_ptr.append(len(term_opt))
                    edge_cursor += 1
            assert edge_cursor - node_edge0[-1] == m * (m - 1)

        i32 = lambda a: np.asarray(a, dtype=np.int32)
        return dict(
            routed_nodes=i32(routed_nodes), routed_edge0=i32(node_edge0), routed_row0=i32(node_row0),
            grp_node=i32(grp_node), grp_up=i32(grp_up), grp_od=i32(grp_od),
            grp_has_virtual=i32(grp_has_virtual), opt_ptr=i32(opt_ptr),
            opt_link=i32(opt_link), opt_slot=i32(opt_slot),
            opt_dist=np.asarray(opt_dist, dtype=np.float64),
            row_ptr=i32(row_ptr), row_od=i32(row_od),
            term_ptr=i32(term_ptr), term_opt=i32(term_opt), term_row_entry=i32(term_row_entry),
        )
