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
        old = getattr(self, "network", None)             # the reference rebuilds the network at every reset (Q10);
        if old is not None and getattr(old, "_engine", None) is not None:
            old._engine.release()                        # ... so hand the old episode's device history back first
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
        return self._observe(), self._info()

    def step(self, actions: Dict[str, Any]):
        unknown = [a for a in actions if a not in self.possible_agents]
        if unknown:
            raise ValueError(f"Unknown agent: {unknown[0]}")
        self.current_actions = actions
        if self.last_actions is None:
            self.last_actions = actions
        if actions:
            self.action_applier.apply_all_actions(actions)
        elif self.sim_step == 1:
            print("No actions provided, skipping action application.")

        earned = dict.fromkeys(self.possible_agents, 0.0)
        result = None
        for _ in range(self._action_gap):                 # `action_gap` simulation steps per decision
            self.network.network_loading(self.sim_step)
            for agent_id, r in self._reward_now().items():
                earned[agent_id] += r
            finished = self.sim_step >= self.simulation_steps      # before the increment: S env steps (Q8)
            result = (self._observe(), dict.fromkeys(self.possible_agents, finished),
                      dict.fromkeys(self.possible_agents, False), self._info())
            self.sim_step += 1
        for agent_id, r in earned.items():
            self._cumulative_rewards[agent_id] += r
        obs, terminations, truncations, infos = result
        return obs, earned, terminations, truncations, infos

    # ------------------------------------------------------------------ per-step quantities
    def _observe(self) -> Dict[str, Any]:
        build = self.obs_builder.build_observation
        return {a: build(a, self.sim_step) for a in self.possible_agents}

    def _gate_reward(self, agent_id: str, t: int) -> float:
        """-sum(T + T_rev) - sum 10 (rho - k_c)[rho > 4] - 10 mean|rho - mean rho| over the agent's links at
        the step just simulated (pz_pednet_env.py:548-581)."""
        total, rho_all = 0.0, []
        for link in self.agent_manager.get_gater_outgoing_links(agent_id):
            back = link.reverse_link
            rho = link.get_density(t)
            rho_all.append(rho)
            t_fwd = link.travel_time[t] if t < len(link.travel_time) else link.travel_time[0]
            t_back = back.travel_time[t] if t < len(back.travel_time) else back.travel_time[0]
            total -= t_fwd + t_back
            if rho > 4:
                total -= 10 * (rho - link.k_critical)
        if len(rho_all) > 1:
            total -= 10.0 * np.mean(np.abs(np.array(rho_all) - np.mean(rho_all)))
        return total

    def _reward_now(self) -> Dict[str, float]:
        """Reference quirk Q2: its loop returns after the first agent, so only that agent can be
        rewarded, and only if it is a gate agent."""
        if not self.possible_agents:
            return {}
        first = self.possible_agents[0]
        if self.agent_manager.get_agent_type(first) != "gate":
            return {}
        return {first: self._gate_reward(first, self.sim_step)}

    def _info(self) -> Dict[str, Dict]:
        return {a: {"step": self.sim_step, "cumulative_reward": self._cumulative_rewards.get(a, 0.0)}
                for a in self.possible_agents}

    def render(self, *a, **k):
        """Nothing to draw without a render mode (reference pz_pednet_env.py:644-648); the plotting itself is
        the reference's NetworkVisualizer, which reads `env.network` or a directory written by `save`."""
        if getattr(self, "render_mode", None) is None:
            return None
        raise NotImplementedError("rendering is outside the accelerated path; use the reference's "
                                  "NetworkVisualizer on env.network or on a directory written by env.save()")

    def save(self, simulation_dir: str, base_dir: str = "../outputs"):
        """Save the current network state (reference pz_pednet_env.py:688-691): the directory layout and JSON
        schema of the reference's OutputHandler, under `base_dir/simulation_dir`."""
        import os
        from ..output import save_network_state
        return save_network_state(self.network, os.path.join(base_dir, simulation_dir))

    def close(self):
        pass
