"""Controller discovery: which links each agent owns (behaviour of the reference's rl/discovery.py:29-178,
whose query names are kept).

Agents, in id order: `sep_{lo}_{hi}` for every corridor listed under `controllers.links` (its two
directions must be Separator links), then `gate_{node}` for every node under `controllers.nodes`; a
gate agent controls the real, non-separator outgoing links of its node in the node's slot order.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

from ..link import Separator


@dataclass
class _Agent:
    kind: str                       # "sep" | "gate"
    links: list = field(default_factory=list)   # sep: [forward, reverse]; gate: controlled outgoing links
    node: object = None             # gate: the gater node
    total_width: float = 0.0        # sep: corridor width shared by the two lanes


def _separator_agent(network, spec: str) -> Tuple[str, _Agent]:
    ends = tuple(int(x) for x in spec.split("-"))
    if len(ends) != 2:
        raise ValueError(f"Separator pair must have exactly 2 nodes: {ends}")
    lo, hi = min(ends), max(ends)
    forward, reverse = network.links.get((lo, hi)), network.links.get((hi, lo))
    if not forward or not reverse:
        raise ValueError(f"Missing bidirectional links for separator {(lo, hi)}")
    if not isinstance(forward, Separator):
        raise ValueError(f"Link {lo}->{hi} is not a Separator. Use Separator links for lane control.")
    return f"sep_{lo}_{hi}", _Agent("sep", [forward, reverse], total_width=forward._width)


def _gate_agent(network, node_id) -> Tuple[str, _Agent]:
    node = network.nodes.get(node_id)
    if node is None:
        raise ValueError(f"Gater node {node_id} not found in network")
    links = [l for l in node.outgoing_links
             if l is not node.virtual_outgoing_link and not isinstance(l, Separator)]
    if not links:
        raise ValueError(f"Gater node {node_id} has no real outgoing links to control")
    return f"gate_{node_id}", _Agent("gate", links, node=node)


class AgentManager:
    def __init__(self, network):
        self.network = network
        self.controller_gaters = network.controller_gaters
        self.controller_separators = network.controller_links
        self._agents: Dict[str, _Agent] = {}
        for spec in self.controller_separators:
            aid, agent = _separator_agent(network, spec)
            self._agents[aid] = agent
        for node_id in self.controller_gaters:
            aid, agent = _gate_agent(network, node_id)
            self._agents[aid] = agent
        self.agent_to_type = {aid: a.kind for aid, a in self._agents.items()}
        self.max_outdegree = max((len(a.links) for a in self._agents.values() if a.kind == "gate"), default=0)

    def _of_kind(self, agent_id: str, kind: str, what: str) -> _Agent:
        agent = self._agents.get(agent_id)
        if agent is None or agent.kind != kind:
            raise ValueError(f"Unknown {what} agent: {agent_id}")
        return agent

    # -- views in the reference's shapes -------------------------------------------------------
    @property
    def separator_agents(self) -> Dict[str, dict]:
        return {aid: {"forward": a.links[0], "reverse": a.links[1], "total_width": a.total_width}
                for aid, a in self._agents.items() if a.kind == "sep"}

    @property
    def gater_agents(self) -> Dict[str, dict]:
        return {aid: {"node": a.node, "out_links": a.links} for aid, a in self._agents.items() if a.kind == "gate"}

    def get_all_agent_ids(self) -> List[str]:
        return list(self._agents)                      # separators first, then gaters (insertion order)

    def get_separator_agents(self):
        return self.separator_agents

    def get_gater_agents(self):
        return self.gater_agents

    def get_agent_type(self, agent_id: str) -> str:
        try:
            return self.agent_to_type[agent_id]
        except KeyError:
            raise ValueError(f"Unknown agent ID: {agent_id}") from None

    def get_separator_links(self, agent_id: str) -> Tuple:
        forward, reverse = self._of_kind(agent_id, "sep", "separator").links
        return forward, reverse

    def get_separator_total_width(self, agent_id: str) -> float:
        return self._of_kind(agent_id, "sep", "separator").total_width

    def get_gater_node(self, agent_id: str):
        return self._of_kind(agent_id, "gate", "gater").node

    def get_gater_outgoing_links(self, agent_id: str) -> List:
        return self._of_kind(agent_id, "gate", "gater").links

    def get_gater_action_mask(self, agent_id: str):
        """float32 [max_outdegree]: 1 for the agent's links, 0 for padding."""
        import numpy as np
        mask = np.zeros(self.max_outdegree, dtype=np.float32)
        mask[: len(self.get_gater_outgoing_links(agent_id))] = 1.0
        return mask

    def get_max_outdegree(self, agent_id: str) -> int:
        return len(self.get_gater_outgoing_links(agent_id))
