"""Host-side node descriptors.

Nodes carry topology (incoming/outgoing link lists in the reference's slot order: the virtual
link first when present, then neighbours by ascending id -- reference: src/LTM/network.py:194-248,
node.py:28-51), the origin demand row and the turning fractions.  The flow assignment itself
(`assign_flows`/`solve`, node.py:164-300) is the node kernel in `csrc/`.
"""
from __future__ import annotations

import numpy as np

from .link import BaseLink


class Node:
    kind = -1

    def __init__(self, node_id, net=None):
        self.node_id = node_id
        self._net = net
        self.incoming_links = []
        self.outgoing_links = []
        self._tf_static = None          # host-owned fractions (default uniform / user supplied)
        self.virtual_incoming_link = None
        self.virtual_outgoing_link = None
        self.M = 1e6                    # receiving flow of a virtual destination link
        self.w = 1e-2                   # weight of the turning-fraction penalty of the 'optimal' model (node.py:14)
        self.demand = None
        self.source_num = None
        self.dest_num = None
        self.edge_num = None
        self.mask = None
        self.q = None
        self.ods_in_turns = {}
        self.index = None               # position in network.nodes (device node index)

    # turning fractions: device-computed for routed nodes, host-owned otherwise
    @property
    def turning_fractions(self):
        if self._net is not None:
            live = self._net._routed_fractions(self)
            if live is not None:
                return live
            if self._tf_static is None and self._net._stepped() and self.edge_num:
                return np.ones(self.edge_num) * (1 / (self.dest_num - 1))     # network.py:269-271
        return self._tf_static

    @turning_fractions.setter
    def turning_fractions(self, value):
        self._tf_static = None if value is None else np.asarray(value, dtype=np.float64)
        if self._net is not None:
            self._net._mark_fractions_dirty()

    def _create_virtual_link(self, store, col, direction, is_incoming):
        link = BaseLink(store, col, f"virtual_{direction}_{self.node_id}",
                        None if is_incoming else self, self if is_incoming else None)
        if is_incoming:
            self.incoming_links.append(link)
            self.virtual_incoming_link = link
        else:
            self.outgoing_links.append(link)
            self.virtual_outgoing_link = link
        return link

    def init_node(self):
        self.source_num = len(self.incoming_links)
        self.dest_num = len(self.outgoing_links)
        self.edge_num = self.dest_num * self.source_num - self.source_num
        self.mask = ~np.eye(self.source_num, dtype=bool)

    def update_matrix_A_eq(self, turning_fractions):
        """External fractions setter used by Network.update_turning_fractions_per_node
        (reference node.py:110-118; the matrices of the 'optimal' node model are built on the device from these
        fractions, csrc/pns_lp.cuh)."""
        tf = np.asarray(turning_fractions, dtype=np.float64)
        assert len(tf) == self.edge_num
        self.turning_fractions = tf


class OneToOneNode(Node):
    kind = 0


class RegularNode(Node):
    kind = 1
