"""Controller discovery: which links each agent owns (reference: rl/discovery.py:29-178).

Agent ids: `sep_{min}_{max}` for every corridor listed in `controllers.links`, then `gate_{node}`
for every node in `controllers.nodes`; a gate agent controls the non-virtual, non-separator
outgoing links of its node in the node's slot order.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from ..link import Separator


class AgentManager:
    def __init__(self, network):
        self.network = network
        self.controller_gaters = network.controller_gaters
        self.controller_separators = network.controller_links
        self.separator_agents: Dict[str, dict] = {}
        self.gater_agents: Dict[str, dict] = {}
        self.agent_to_type: Dict[str, str] = {}
        self._find_separators()
        self._find_gaters()
        self.max_outdegree = max((len(a["out_links"]) for a in self.gater_agents.values()), default=0)

    def _find_separators(self):
        for spec in self.controller_separators:
            pair = tuple(int(x) for x in spec.split("-"))
            if len(pair) != 2:
                raise ValueError(f"Separator pair must have exactly 2 nodes: {pair}")
            lo, hi = sorted(pair)
            fwd, rev = self.network.links.get((lo, hi)), self.network.links.get((hi, lo))
            if not fwd or not rev:
                raise ValueError(f"Missing bidirectional links for separator {(lo, hi)}")
            if not isinstance(fwd, Separator):
                raise ValueError(f"Link {lo}->{hi} is not a Separator. Use Separator links for lane control.")
            aid = f"sep_{lo}_{hi}"
            self.separator_agents[aid] = {"forward": fwd, "reverse": rev, "total_width": fwd._width}
            self.agent_to_type[aid] = "sep"

    def _find_gaters(self):
        for node_id in self.controller_gaters:
            if node_id not in self.network.nodes:
                raise ValueError(f"Gater node {node_id} not found in network")
            node = self.network.nodes[node_id]
            owned = [l for l in node.outgoing_links
                     if not isinstance(l, Separator) and l is not node.virtual_outgoing_link]
            if not owned:
                raise ValueError(f"Gater node {node_id} has no real outgoing links to control")
            aid = f"gate_{node_id}"
            self.gater_agents[aid] = {"node": node, "out_links": owned}
            self.agent_to_type[aid] = "gate"

    # -- queries (same names as the reference) -------------------------------------------------
    def get_all_agent_ids(self) -> List[str]:
        return list(self.separator_agents) + list(self.gater_agents)

    def get_separator_agents(self):
        return dict(self.separator_agents)

    def get_gater_agents(self):
        return dict(self.gater_agents)

    def get_agent_type(self, agent_id: str) -> str:
        if agent_id not in self.agent_to_type:
            raise ValueError(f"Unknown agent ID: {agent_id}")
        return self.agent_to_type[agent_id]

    def get_separator_links(self, agent_id: str) -> Tuple:
        if agent_id not in self.separator_agents:
            raise ValueError(f"Unknown separator agent: {agent_id}")
        a = self.separator_agents[agent_id]
        return a["forward"], a["reverse"]

    def get_separator_total_width(self, agent_id: str) -> float:
        if agent_id not in self.separator_agents:
            raise ValueError(f"Unknown separator agent: {agent_id}")
        return self.separator_agents[agent_id]["total_width"]

    def get_gater_node(self, agent_id: str):
        if agent_id not in self.gater_agents:
            raise ValueError(f"Unknown gater agent: {agent_id}")
        return self.gater_agents[agent_id]["node"]

    def get_gater_outgoing_links(self, agent_id: str) -> List:
        if agent_id not in self.gater_agents:
            raise ValueError(f"Unknown gater agent: {agent_id}")
        return self.gater_agents[agent_id]["out_links"]

    def get_gater_action_mask(self, agent_id: str) -> np.ndarray:
        n = len(self.get_gater_outgoing_links(agent_id))
        mask = np.zeros(self.max_outdegree, dtype=np.float32)
        mask[:n] = 1.0
        return mask

    def get_max_outdegree(self, agent_id: str) -> int:
        return len(self.get_gater_outgoing_links(agent_id))
