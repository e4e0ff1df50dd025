"""The stub a maintainer of the reference would add: `B200Step(network)` attaches the CUDA timestep to the
reference's *own* `src.LTM.network.Network` object and replaces the body of `network_loading`
(reference src/LTM/network.py:266-287).

    from pednstream_b200.reference_adapter import B200Step
    net = NetworkEnvGenerator().create_network("nine_intersections")      # the reference's classes, untouched
    b200 = B200Step(net)                                                   # plan + device state from its objects
    for t in range(1, net.simulation_steps):
        b200.step(t)            # == net.network_loading(t); every per-link array of `net` is filled in

With `assign_flows_type: optimal` the device solves each node's linear program itself (csrc/pns_lp.cuh): the flows
are optimal solutions of the same programs, but where an optimum is not unique, or a flow is an integer up to
rounding, the vertex -- and so the trajectory -- is not the one scipy's solver would have produced.

The reference object stays the source of truth for its consumers (output handler, visualiser, RL builders):
after every step the rows the step produced are written into the reference's own numpy arrays
(`link.inflow[t]`, ..., `link.sending_flow[t-1]`), and host mutations made on the reference objects between
steps -- gate and separator widths, demand entries, OD weights (examples/long_corridor.py:65-66, 124-134) --
are read from them before the step.  Topology, parameters, demand, OD weights, k-shortest paths and the logit
offset are taken from the reference object; nothing is drawn from numpy's stream while attaching.  In the default
numpy-compatible draw mode the in-step random numbers come from `np.random` in the reference's visiting
order, so `b200.step(t)` leaves the reference object in the state its own `network_loading(t)` would.
"""
from __future__ import annotations

import copy

import numpy as np

from .network import Network
from .state import F32_FIELDS, F64_FIELDS

_LINK_FIELDS = F64_FIELDS[:7] + F32_FIELDS


class B200Step:
    def __init__(self, network, rng: str = "numpy", seed: int = 0, device=None, _lib=None, _emulation: bool = False):
        self.network = network
        od = getattr(network, "od_manager", None)
        od_flows = {k: np.array(v, dtype=np.float64) for k, v in od.od_flows.items()} if od is not None else None
        state = np.random.get_state()               # attaching must not move the reference's global stream
        try:
            params = copy.deepcopy(network.params)
            params["assign_flows_type"] = getattr(network, "assign_flows_type", "classic")
            self.facade = Network(network.adjacency_matrix, params,
                                  list(network.origin_nodes), list(network.destination_nodes),
                                  od_flows=od_flows, pos=getattr(network, "pos", None), verbose=False,
                                  rng=rng, seed=seed, device=device, _lib=_lib, _emulation=_emulation)
        finally:
            np.random.set_state(state)
        f = self.facade
        if list(f.links.keys()) != list(network.links.keys()) or list(f.nodes.keys()) != list(network.nodes.keys()):
            raise RuntimeError("link / node order differs from the reference object")
        for nid, node in network.nodes.items():      # the reference's demand draws, not new ones
            if node.demand is not None:
                f.nodes[nid].demand = np.array(node.demand)
        if getattr(network, "path_finder", None) is not None:
            ref_paths = {k: [list(map(int, p)) for p in v] for k, v in network.path_finder.od_paths.items()}
            own_paths = {k: [list(map(int, p)) for p in v] for k, v in f.path_finder.od_paths.items()}
            if ref_paths != own_paths:
                raise RuntimeError("k-shortest paths differ from the reference object's")
            f.path_finder.epsilon = network.path_finder.epsilon
        self._pairs = [(network.links[k], f.links[k]) for k in network.links]
        self._virtual = []
        for nid, node in network.nodes.items():
            fn = f.nodes[nid]
            if getattr(node, "virtual_incoming_link", None) is not None:
                self._virtual.append((node.virtual_incoming_link, fn.virtual_incoming_link))
                self._virtual.append((node.virtual_outgoing_link, fn.virtual_outgoing_link))
        self._demand_nodes = [(node, f.nodes[nid]) for nid, node in network.nodes.items() if node.demand is not None]
        self._tf_seen = {}

    # ------------------------------------------------------------------ host edits made on the reference objects
    def _pull_edits(self, t):
        for ref, mine in self._pairs:
            if mine.is_separator:
                w = ref.separator_width
                if w != mine.separator_width or type(w) is not type(mine.separator_width):
                    mine.separator_width = w
            elif ref.back_gate_width != mine.back_gate_width:
                mine.back_gate_width = ref.back_gate_width
        for ref, mine in self._demand_nodes:
            if t - 1 < len(ref.demand) and mine.demand[t - 1] != ref.demand[t - 1]:
                mine.demand[t - 1] = ref.demand[t - 1]
        od = getattr(self.network, "od_manager", None)
        if od is not None:
            for k, series in od.od_flows.items():
                mine = self.facade.od_manager.od_flows[k]
                if t < len(series) and mine[t] != series[t]:
                    mine[t] = series[t]
        pf = getattr(self.network, "path_finder", None)
        for nid, node in self.network.nodes.items():          # fractions set from outside (network.py:250-255)
            routed = pf is not None and nid in pf.nodes_in_paths and node.source_num > 2
            tf = node.turning_fractions
            if routed or tf is None:
                continue
            key = tf.tobytes() if hasattr(tf, "tobytes") else bytes(np.asarray(tf))
            if self._tf_seen.get(nid) != key:
                self._tf_seen[nid] = key
                self.facade.nodes[nid].turning_fractions = np.array(tf, dtype=np.float64)

    # ------------------------------------------------------------------ results into the reference's arrays
    def _push_rows(self, t):
        store = self.facade._store
        L = len(self._pairs)
        rows = {f: store.field(f) for f in _LINK_FIELDS}
        for ref, mine in self._pairs:
            c = mine._col
            for f in _LINK_FIELDS:
                a = getattr(ref, f)
                a[t] = rows[f][t, c]
            ref.sending_flow[t - 1] = rows["sending_flow"][t - 1, c]
            ref.receiving_flow[t - 1] = rows["receiving_flow"][t - 1, c]
            if mine.is_separator:
                ref.separator_width_data[t] = store.field("separator_width_data")[t, c]
        for ref, mine in self._virtual:
            c = mine._col
            for f in ("inflow", "outflow", "cumulative_inflow", "cumulative_outflow"):
                getattr(ref, f)[t] = rows[f][t, c]
        pf = getattr(self.network, "path_finder", None)
        for nid, node in self.network.nodes.items():
            if pf is not None and nid in pf.nodes_in_paths and node.source_num > 2:
                node.turning_fractions = np.array(self.facade.nodes[nid].turning_fractions)
            elif node.turning_fractions is None:                      # network.py:269-271
                node.turning_fractions = np.ones(node.edge_num) * (1 / (node.dest_num - 1))

    def step(self, t: int):
        """`network.network_loading(t)` on the device."""
        self._pull_edits(int(t))
        self.facade.network_loading(int(t))
        self._push_rows(int(t))

    def install(self):
        """Rebind `network.network_loading` to the device step (what the one-line change in
        src/LTM/network.py:266 amounts to)."""
        self.network.network_loading = self.step
        return self
