"""Observation layout and action clipping for the single-network environment
(reference: rl/builders.py:25-352).  Layout is link-major with the gate width as the last feature
of every link block; trainers recover the current width from it (rl/rl_utils.py:1575-1577).

The same layout and clipping rules are implemented on the device for the batched environment
(csrc/pns_kernels.cu: k_env_actions, k_env_observe) -- `OBS_LAYOUT` is shared by both.
"""
from __future__ import annotations

import numpy as np

# features per controlled out-link, in order.  'rev.' = the reverse link; 'gdens' = shared density
OBS_LAYOUT = {
    "option1": ("inflow", "rev.outflow", "gate"),
    "option2": ("inflow", "rev.outflow", "gdens", "gate"),
    "option3": ("inflow", "outflow", "rev.inflow", "rev.outflow", "gate"),
    "option4": ("gdens/kjam", "gate"),
    "option5": ("inflow", "outflow", "rev.inflow", "rev.outflow", "speed", "gdens", "gate"),
}
DENSITY_NORM, FLOW_NORM = 6.0, 20.0


def _at(series, t):
    return series[t] if t < len(series) else 0.0


class ObservationBuilder:
    def __init__(self, network, agent_manager, normalize: bool = False, obs_mode: str = "flow_only"):
        if obs_mode not in OBS_LAYOUT:
            raise ValueError(f"obs_mode must be one of {list(OBS_LAYOUT)}, got: {obs_mode}")
        self.network, self.agent_manager = network, agent_manager
        self.normalize, self.obs_mode = normalize, obs_mode
        self.features_per_link = len(OBS_LAYOUT[obs_mode])
        self.density_norm, self.speed_norm, self.flow_norm = DENSITY_NORM, 1.5, FLOW_NORM

    def build_observation(self, agent_id: str, time_step: int) -> np.ndarray:
        kind = self.agent_manager.get_agent_type(agent_id)
        if kind == "sep":
            return self._separator_obs(agent_id, time_step)
        if kind == "gate":
            return self._gater_obs(agent_id, time_step)
        raise ValueError(f"Unknown agent type: {kind}")

    def _feature(self, link, name, t):
        if name == "gate":
            return link.back_gate_width
        if name == "gdens":
            return link.get_density(t)
        if name == "gdens/kjam":
            return link.get_density(t) / link.k_jam if t < len(link.speed) else 0.0
        src = link
        if name.startswith("rev."):
            src, name = link.reverse_link, name[4:]
        return _at(getattr(src, name), t)

    def _gater_obs(self, agent_id, t):
        links = self.agent_manager.get_gater_outgoing_links(agent_id)
        k = self.features_per_link
        obs = np.zeros(self.agent_manager.get_max_outdegree(agent_id) * k, dtype=np.float32)
        for i, link in enumerate(links):
            obs[i * k:(i + 1) * k] = [self._feature(link, f, t) for f in OBS_LAYOUT[self.obs_mode]]
        return self._normalize_gater_obs(obs) if self.normalize else obs

    def _separator_obs(self, agent_id, t):
        fwd, rev = self.agent_manager.get_separator_links(agent_id)
        obs = np.array([_at(fwd.inflow, t), _at(fwd.outflow, t), _at(rev.inflow, t), _at(rev.outflow, t)],
                       dtype=np.float32)
        return self._normalize_separator_obs(obs) if self.normalize else obs

    # fixed-constant normalisation, index rules exactly as rl/builders.py:179-238 (including the
    # indices that fall outside short vectors, which raise IndexError there as well)
    def _normalize_gater_obs(self, obs):
        out = obs.copy()
        k = self.features_per_link
        if k == 0:
            return out
        for i in range(len(obs) // k):
            s = i * k
            if self.obs_mode in ("option1", "option2"):
                out[s] /= self.flow_norm
                out[s + 1] /= self.flow_norm
            elif self.obs_mode in ("option3", "option4"):
                out[s] /= self.density_norm
                out[s + 1] /= self.flow_norm
                out[s + 2] /= self.flow_norm
        return out

    def _normalize_separator_obs(self, obs):
        out = obs.copy()
        if self.obs_mode == "option1":
            out[:] /= self.flow_norm
        elif self.obs_mode == "option2":
            out[:4] /= self.flow_norm
        elif self.obs_mode in ("option3", "option4"):
            out[[0, 3]] /= self.density_norm
            out[[1, 2, 4, 5]] /= self.flow_norm
        return out


class ActionApplier:
    """Actions are absolute widths in metres; rate-limited per env step, then clipped."""

    def __init__(self, network, agent_manager, max_delta_sep_width: float = 0.1,
                 max_delta_gate_width: float = 0.1, min_sep_width: float = 1.0):
        self.network, self.agent_manager = network, agent_manager
        self.max_delta_sep_width = max_delta_sep_width
        self.max_delta_gate_width = max_delta_gate_width
        self.min_sep_width = min_sep_width

    def apply_all_actions(self, actions):
        for agent_id, action in actions.items():
            kind = self.agent_manager.get_agent_type(agent_id)
            if kind == "sep":
                fwd, _ = self.agent_manager.get_separator_links(agent_id)
                fwd.separator_width = self.clip_separator_action_value(float(action[0]), fwd)
            elif kind == "gate":
                for i, link in enumerate(self.agent_manager.get_gater_outgoing_links(agent_id)):
                    link.back_gate_width = self.clip_gater_action_value(float(action[i]), link)
            else:
                raise ValueError(f"Unknown agent type: {kind}")

    @staticmethod
    def _rate_limit(value, current, max_delta):
        if abs(value - current) > max_delta:
            value = current + np.clip(value - current, -max_delta, max_delta)
        return value

    def clip_separator_action_value(self, action_value, forward_link):
        v = self._rate_limit(action_value, forward_link.separator_width, self.max_delta_sep_width)
        return np.clip(v, self.min_sep_width, forward_link.width - self.min_sep_width)

    def clip_gater_action_value(self, action_value, link):
        v = self._rate_limit(action_value, link.back_gate_width, self.max_delta_gate_width)
        return np.clip(v, 0.0, link.width)
