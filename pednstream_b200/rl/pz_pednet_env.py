"""`PedNetParallelEnv`: the reference's PettingZoo-style control environment for one network
(reference: rl/pz_pednet_env.py:38-254, 548-642), on top of the B200 timestep.

Same constructor, `reset`/`step`/`seed`/`agents`/`possible_agents`/`observation_space`/
`action_space`/`close`; observations, rewards, terminations, truncations and infos are dicts keyed
by agent id.  pettingzoo / gymnasium are optional (duck-typed base class and spaces).

Reference behaviours kept on purpose (SURVEY.md section 8b "quirks"): the `seed` argument of
`reset` is ignored (Q5); only the first agent in `possible_agents` order receives a reward if it is
a gate agent, because the reference returns from inside its loop (Q2, pz_pednet_env.py:581);
termination is tested before `sim_step` is incremented, so an episode is exactly S env steps (Q8).
"""
from __future__ import annotations

import random
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from ..env_loader import NetworkEnvGenerator
from .builders import ActionApplier, ObservationBuilder
from .discovery import AgentManager
from .spaces import SpaceBuilder

try:                                    # pragma: no cover - optional dependency
    from pettingzoo import ParallelEnv as _Base
except Exception:
    class _Base:                        # minimal stand-in
        pass


class PedNetParallelEnv(_Base):
    metadata = {"render_modes": ["human", "animate"], "name": "pednet_v0"}

    def __init__(self, dataset: str, normalize_obs: bool = False, obs_mode: str = "option1",
                 render_mode: Optional[str] = None, verbose: bool = False, action_gap: int = 1,
                 seed: Optional[int] = None, **engine_kw):
        """engine_kw (rng=, device=) is forwarded to the network; `rng="numpy"` (default) reproduces
        the reference's trajectories for the same seed."""
        self.render_mode = render_mode
        self.verbose = verbose
        self._engine_kw = engine_kw
        self._seed = seed
        if seed is not None:
            np.random.seed(seed)
            random.seed(seed)
        self.env_generator = NetworkEnvGenerator()
        self.dataset = dataset
        self.network = self.env_generator.create_network(dataset, verbose=verbose, **engine_kw)
        self.sim_step = 1
        self.simulation_steps = self.network.params["simulation_steps"]
        self._max_delta_sep_width = 0.25 * self.network.params["unit_time"]
        self._max_delta_gate_width = 0.25 * self.network.params["unit_time"]
        self._min_sep_width = 1.5
        self.normalize_obs = normalize_obs
        self.obs_mode = obs_mode
        self._bind_network()
        self.possible_agents = self.agent_manager.get_all_agent_ids()
        self.space_builder = SpaceBuilder(self.agent_manager, self.obs_mode, self._min_sep_width)
        self._action_spaces = self.space_builder.build_action_spaces()
        self._observation_spaces = self.space_builder.build_observation_spaces(self.obs_builder.features_per_link)
        self._cumulative_rewards = {a: 0.0 for a in self.possible_agents}
        self._action_gap = action_gap
        self.last_actions = None
        self.current_actions = None
        self.visualizer = None

    def _bind_network(self):
        self.agent_manager = AgentManager(self.network)
        self.obs_builder = ObservationBuilder(self.network, self.agent_manager, self.normalize_obs, self.obs_mode)
        self.action_applier = ActionApplier(self.network, self.agent_manager, self._max_delta_sep_width,
                                            self._max_delta_gate_width, self._min_sep_width)

    def seed(self, seed: int) -> None:
        self._seed = seed
        np.random.seed(seed)
        random.seed(seed)

    @property
    def agents(self) -> List[str]:
        return self.possible_agents.copy()

    def observation_space(self, agent: str):
        if agent not in self._observation_spaces:
            raise ValueError(f"Agent {agent} not found in observation spaces")
        return self._observation_spaces[agent]

    def action_space(self, agent: str):
        if agent not in self._action_spaces:
            raise ValueError(f"Agent {agent} not found in action spaces")
        return self._action_spaces[agent]

    # ------------------------------------------------------------------ episode control
    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None) -> Tuple[Dict, Dict]:
        if options and options.get("randomize", False):
            # the reference passes `verbose=` to a randomize_network that does not take it and raises
            # TypeError (pz_pednet_env.py:163-165); this is the evident intent of that branch
            self.network = self.env_generator.randomize_network(self.dataset, seed=None, verbose=self.verbose,
                                                                **self._engine_kw)
        else:
            self.network = self.env_generator.create_network(self.dataset, verbose=self.verbose,
                                                             **self._engine_kw)
        self._bind_network()
        self.sim_step = 1
        self._cumulative_rewards = {a: 0.0 for a in self.possible_agents}
        return self._get_observations(), self._get_infos()

    def step(self, actions: Dict[str, Any]):
        self.current_actions = actions
        if self.last_actions is None:
            self.last_actions = actions
        for agent_id in actions:
            if agent_id not in self.possible_agents:
                raise ValueError(f"Unknown agent: {agent_id}")
        if len(actions) > 0:
            self.action_applier.apply_all_actions(actions)
        elif self.sim_step == 1:
            print("No actions provided, skipping action application.")

        gap_rewards = {a: 0.0 for a in self.possible_agents}
        observations = terminations = truncations = infos = None
        for _ in range(self._action_gap):
            self.network.network_loading(self.sim_step)
            observations = self._get_observations()
            for agent_id, r in self._compute_rewards().items():
                gap_rewards[agent_id] += r
            terminations = self._check_terminations()
            truncations = {a: False for a in self.possible_agents}
            infos = self._get_infos()
            self.sim_step += 1
        for agent_id, r in gap_rewards.items():
            self._cumulative_rewards[agent_id] += r
        return observations, gap_rewards, terminations, truncations, infos

    # ------------------------------------------------------------------ per-step quantities
    def _get_observations(self) -> Dict[str, Any]:
        return {a: self.obs_builder.build_observation(a, self.sim_step) for a in self.possible_agents}

    def _compute_rewards(self) -> Dict[str, float]:
        """Gate agent: -sum(T + T_rev) - sum 10 (rho - k_c)[rho > 4] - 10 mean|rho - mean rho|
        over its controlled links at the step just simulated (pz_pednet_env.py:548-581)."""
        rewards = {}
        t = self.sim_step
        for agent_id in self.possible_agents:
            if self.agent_manager.get_agent_type(agent_id) == "gate":
                total = 0.0
                densities = []
                for link in self.agent_manager.get_gater_outgoing_links(agent_id):
                    rho = link.get_density(t)
                    densities.append(rho)
                    rev = link.reverse_link
                    T = link.travel_time[t] if t < len(link.travel_time) else link.travel_time[0]
                    T_rev = rev.travel_time[t] if t < len(rev.travel_time) else rev.travel_time[0]
                    total -= T + T_rev
                    if rho > 4:
                        total -= 10 * (rho - link.k_critical)
                if len(densities) > 1:
                    mean = np.mean(densities)
                    total -= 10.0 * np.mean(np.abs(np.array(densities) - mean))
                rewards[agent_id] = total
            return rewards          # reference quirk Q2: only the first agent is considered
        return rewards

    def _check_terminations(self) -> Dict[str, bool]:
        done = self.sim_step >= self.simulation_steps
        return {a: done for a in self.possible_agents}

    def _get_infos(self) -> Dict[str, Dict]:
        return {a: {"step": self.sim_step, "cumulative_reward": self._cumulative_rewards.get(a, 0.0)}
                for a in self.possible_agents}

    def render(self, *a, **k):
        raise NotImplementedError("rendering is outside the accelerated path; use the reference's "
                                  "NetworkVisualizer on env.network")

    def save(self, simulation_dir: str):
        raise NotImplementedError("use the reference's OutputHandler on env.network")

    def close(self):
        pass
